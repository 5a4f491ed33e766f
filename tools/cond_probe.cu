// What does an IF node whose condition is false cost inside a CUDA graph on B200?  (lost-object mode: every step carries one.)
//   chain A: N x [k_work -> k_work -> k_work]                     (three plain kernel nodes per step)
//   chain B: N x [k_work -> k_mark(sets cond = 0) -> IF{k_work} -> k_work]
//   chain C: N x [k_work(sets cond = 0) -> IF{k_work}]             (the update sets the condition itself; nothing behind the IF)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cond_probe tools/cond_probe.cu ; prints us per step for each chain
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void k_work(unsigned long long* p) { if (threadIdx.x == 0) *p += 1ull; }
__global__ void k_mark(unsigned long long* p, cudaGraphConditionalHandle h, unsigned int v)
{
    if (threadIdx.x == 0) { *p += 1ull; cudaGraphSetConditional(h, v); }
}

static int add_kernel(cudaGraph_t g, cudaGraphNode_t* out, const cudaGraphNode_t* dep, void* fn, void** args)
{
    cudaKernelNodeParams kp{};
    kp.func = fn; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = args;
    CK(cudaGraphAddKernelNode(out, g, dep, dep ? 1 : 0, &kp));
    return 0;
}

int main()
{
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 8));
    CK(cudaMemset(d, 0, 8));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int N = 20;
    for (int variant = 0; variant < 3; ++variant) {
        cudaGraph_t g;
        CK(cudaGraphCreate(&g, 0));
        cudaGraphNode_t prev{}; bool have = false;
        for (int s = 0; s < N; ++s) {
            void* wa[1] = {&d};
            cudaGraphNode_t n1, n2, n3, nif;
            if (variant == 0) {
                if (add_kernel(g, &n1, have ? &prev : nullptr, (void*)k_work, wa)) return 1;
                if (add_kernel(g, &n2, &n1, (void*)k_work, wa)) return 1;
                if (add_kernel(g, &n3, &n2, (void*)k_work, wa)) return 1;
                prev = n3;
            } else {
                cudaGraphConditionalHandle h;
                CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
                unsigned int v = 0;
                void* ma[3] = {&d, &h, &v};
                if (variant == 1) {
                    if (add_kernel(g, &n1, have ? &prev : nullptr, (void*)k_work, wa)) return 1;
                    if (add_kernel(g, &n2, &n1, (void*)k_mark, ma)) return 1;
                } else {
                    if (add_kernel(g, &n2, have ? &prev : nullptr, (void*)k_mark, ma)) return 1;
                }
                cudaGraphNodeParams cp{};
                cp.type = cudaGraphNodeTypeConditional;
                cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
                CK(cudaGraphAddNode(&nif, g, &n2, 1, &cp));
                cudaGraph_t body = cp.conditional.phGraph_out[0];
                cudaGraphNode_t nb;
                if (add_kernel(body, &nb, nullptr, (void*)k_work, wa)) return 1;
                if (variant == 1) { if (add_kernel(g, &n3, &nif, (void*)k_work, wa)) return 1; prev = n3; }
                else prev = nif;
            }
            have = true;
        }
        cudaGraphExec_t ge;
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphUpload(ge, st));
        for (int w = 0; w < 5; ++w) CK(cudaGraphLaunch(ge, st));
        CK(cudaStreamSynchronize(st));
        const int R = 50;
        CK(cudaEventRecord(a, st));
        for (int r = 0; r < R; ++r) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(b, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        const char* what[3] = {"A: 3 kernels per step", "B: kernel, mark, IF(false), kernel", "C: mark, IF(false)"};
        printf("%s: %.2f us per step\n", what[variant], 1e3 * ms / (R * N));
        CK(cudaGraphExecDestroy(ge));
        CK(cudaGraphDestroy(g));
    }
    return 0;
}
