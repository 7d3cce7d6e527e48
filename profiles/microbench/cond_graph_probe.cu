// Probe: nested WHILE conditional graph nodes filled by stream capture, kernels launched with the programmatic
// dependent launch attribute inside the loop bodies.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void work(int *cnt, int slot) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) cnt[slot] += 1;
}
__global__ void inner_decide(cudaGraphConditionalHandle h, int *cnt) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int c = ++cnt[2];                      // inner trips of this outer iteration
    int go = c < 1 + cnt[3];               // outer iteration k runs k+1 inner trips
    if (!go) cnt[2] = 0;
    cudaGraphSetConditional(h, go);
}
__global__ void outer_decide(cudaGraphConditionalHandle h, int *cnt) {
    int c = ++cnt[3];
    cudaGraphSetConditional(h, c < 4);
}

template <class... A>
cudaError_t launch(void (*k)(A...), cudaStream_t s, bool pdl, A... a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4); cfg.blockDim = dim3(64); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, a...);
}

static int add_while(cudaStream_t s, cudaGraphConditionalHandle *h, cudaGraph_t *body) {
    cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t g; const cudaGraphNode_t *deps; size_t nd;
    CK(cudaStreamGetCaptureInfo_v2(s, &st, &id, &g, &deps, &nd));
    CK(cudaGraphConditionalHandleCreate(h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = *h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, deps, nd, &p));
    *body = p.conditional.phGraph_out[0];
    CK(cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies));
    return 0;
}

int main(int argc, char **argv) {
    bool pdl = argc > 1;
    cudaStream_t s, s2, s3; CK(cudaStreamCreate(&s)); CK(cudaStreamCreate(&s2)); CK(cudaStreamCreate(&s3));
    int *cnt; CK(cudaMalloc(&cnt, 64)); CK(cudaMemset(cnt, 0, 64));
    cudaGraph_t g, outer_body, inner_body;
    cudaGraphConditionalHandle ho, hi;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    CK(launch(work, s, false, cnt, 0));
    if (add_while(s, &ho, &outer_body)) return 1;
    CK(cudaStreamBeginCaptureToGraph(s2, outer_body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    CK(launch(work, s2, false, cnt, 1));
    CK(launch(work, s2, pdl, cnt, 1));
    if (add_while(s2, &hi, &inner_body)) return 1;
    CK(cudaStreamBeginCaptureToGraph(s3, inner_body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    CK(launch(work, s3, false, cnt, 4));
    CK(launch(work, s3, pdl, cnt, 4));
    CK(launch(inner_decide, s3, pdl, hi, cnt));
    CK(cudaStreamEndCapture(s3, nullptr));
    CK(launch(outer_decide, s2, false, ho, cnt));
    CK(cudaStreamEndCapture(s2, nullptr));
    CK(launch(work, s, false, cnt, 5));
    CK(cudaStreamEndCapture(s, &g));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(cnt, 0, 64));
        cudaEventRecord(a, s);
        CK(cudaGraphLaunch(ex, s));
        cudaEventRecord(b, s);
        CK(cudaStreamSynchronize(s));
        float ms; cudaEventElapsedTime(&ms, a, b);
        int h[8]; CK(cudaMemcpy(h, cnt, 32, cudaMemcpyDeviceToHost));
        // expect pre 1, outer work 8 (4 iterations x 2), inner work 2*(1+2+3+4)=20, outer trips 4, post 1
        printf("pdl=%d pre %d outer_work %d inner_left %d outer_trips %d inner_work %d post %d  %.1f us (%d kernels)\n", (int)pdl, h[0], h[1],
               h[2], h[3], h[4], h[5], ms * 1e3, 2 + 8 + 4 + 20 + 10);
    }
    return 0;
}
