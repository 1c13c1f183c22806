// launch-gap microbenchmark: chains of short dependent kernels issued as plain launches, graphs, PDL launches
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(err_)); return 1; } } while (0)
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void spin(unsigned long long ns, int pdl, unsigned long long* sink) {
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned long long t0 = gt();
    while (gt() - t0 < ns) { }
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = t0;
}
static cudaError_t launch(cudaStream_t s, int grid, unsigned long long ns, int pdl, unsigned long long* sink) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, spin, ns, pdl, sink);
}
int main() {
    cudaStream_t s, s2; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, evs[4]; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (auto& e : evs) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    const unsigned long long NS = 10000; const int K = 4, N = 200;
    for (int grid : {1, 150}) {
        auto report = [&](const char* name, float ms) {
            printf("grid %3d %-44s per-kernel period %.2f us (gap %.2f us)\n", grid, name, ms * 1e3 / (N * K), ms * 1e3 / (N * K) - NS / 1e3);
        };
        float ms;
        for (int pdl = 0; pdl < 2; ++pdl) {
            // 1. plain launches
            CK(launch(s, grid, NS, 0, nullptr)); CK(cudaStreamSynchronize(s));
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N * K; ++i) CK(launch(s, grid, NS, pdl, nullptr));
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            report(pdl ? "plain launches + PDL" : "plain launches", ms);
            // 2. graph of K kernels, launched N times
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < K; ++i) CK(launch(s, grid, NS, pdl, nullptr));
            CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0)); CK(cudaGraphUpload(ge, s));
            CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N; ++i) CK(cudaGraphLaunch(ge, s));
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            report(pdl ? "graph(4) x N + PDL inside" : "graph(4) x N", ms);
            // 3. same with 3 already-signalled event waits + 2 records between graph launches
            for (auto& e : evs) CK(cudaEventRecord(e, s2));
            CK(cudaStreamSynchronize(s2));
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N; ++i) {
                for (int q = 0; q < 3; ++q) CK(cudaStreamWaitEvent(s, evs[q], 0));
                CK(cudaGraphLaunch(ge, s));
                CK(cudaEventRecord(evs[3], s));
            }
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            report(pdl ? "graph(4) x N + PDL, 3 waits + 1 record" : "graph(4) x N, 3 waits + 1 record", ms);
            // 4. first kernel plain, rest graph(3)
            cudaGraph_t g3; cudaGraphExec_t ge3;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < K - 1; ++i) CK(launch(s, grid, NS, pdl, nullptr));
            CK(cudaStreamEndCapture(s, &g3)); CK(cudaGraphInstantiate(&ge3, g3, 0)); CK(cudaGraphUpload(ge3, s));
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N; ++i) { CK(launch(s, grid, NS, pdl, nullptr)); CK(cudaGraphLaunch(ge3, s)); }
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            report(pdl ? "plain(1) + graph(3), PDL" : "plain(1) + graph(3)", ms);
            // 5. one graph holding 6 frames (24 kernels)
            cudaGraph_t g6; cudaGraphExec_t ge6;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < 6 * K; ++i) CK(launch(s, grid, NS, pdl, nullptr));
            CK(cudaStreamEndCapture(s, &g6)); CK(cudaGraphInstantiate(&ge6, g6, 0)); CK(cudaGraphUpload(ge6, s));
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N / 6; ++i) CK(cudaGraphLaunch(ge6, s));
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("grid %3d %-44s per-kernel period %.2f us (gap %.2f us)\n", grid, pdl ? "graph(24) + PDL" : "graph(24)", ms * 1e3 / ((N / 6) * 6 * K), ms * 1e3 / ((N / 6) * 6 * K) - NS / 1e3);
            // 7. plain launches, every 4th followed by 2 records + 3 (signalled) waits -- the frame boundary of the real pipeline
            for (auto& e : evs) CK(cudaEventRecord(e, s2));
            CK(cudaStreamSynchronize(s2));
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N; ++i) {
                for (int q = 0; q < 3; ++q) CK(cudaStreamWaitEvent(s, evs[q], 0));
                for (int k = 0; k < K; ++k) CK(launch(s, grid, NS, pdl, nullptr));
                CK(cudaEventRecord(evs[3], s)); CK(cudaEventRecord(e1, s));
            }
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            report(pdl ? "plain + PDL, 3 waits + 2 records per 4" : "plain, 3 waits + 2 records per 4", ms);
            // 6. cross-stream event dependency: kernel on s, record, s2 waits, kernel on s2, record, s waits ...
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < N * K / 2; ++i) {
                CK(launch(s, grid, NS, 0, nullptr)); CK(cudaEventRecord(evs[0], s)); CK(cudaStreamWaitEvent(s2, evs[0], 0));
                CK(launch(s2, grid, NS, 0, nullptr)); CK(cudaEventRecord(evs[1], s2)); CK(cudaStreamWaitEvent(s, evs[1], 0));
            }
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
            if (!pdl) report("ping-pong across two streams (events)", ms);
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaGraphExecDestroy(ge3); cudaGraphDestroy(g3); cudaGraphExecDestroy(ge6); cudaGraphDestroy(g6);
        }
    }
    return 0;
}
