// Development probe (GPU): what do the two SMs of a TPC share?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tpc_probe scripts/tpc_probe.cu && scripts/tpc_probe
//
// Two kernels, one CTA of 128 threads per SM, timed per CTA with clock64 for grids that occupy one SM per TPC (<= 74 CTAs)
// and both (148):
//   loop      8 independent DFMA chains per thread in a loop that fits the L0 instruction cache (FP64 pipe only)
//   straight  the same DFMAs as ~112 KB of straight-line code behind an outer loop (instruction delivery + FP64 pipe)
// The backward sweeps are ~110 KB of straight-line FP64 code per stage; if `straight` slows down when both SMs of a TPC run
// it and `loop` does not, the sweeps are bound by instruction delivery shared inside the TPC.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

template <int BODY>
__global__ void __launch_bounds__(128, 1) k_probe(int outer, double c, double* out, long long* cyc, unsigned* sm) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int k = 0; k < BODY; ++k) x[k & 7] = fma(x[k & 7], c, x[(k + 3) & 7]);
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) {
        cyc[blockIdx.x] = t1 - t0;
        sm[blockIdx.x] = smid();
    }
}

template <int BODY>
void run(const char* name, int grid, long long total) {
    const int outer = (int)(total / BODY);
    double* out;
    long long* cyc;
    unsigned* sm;
    cudaMalloc(&out, grid * 128 * sizeof(double));
    cudaMalloc(&cyc, grid * sizeof(long long));
    cudaMalloc(&sm, grid * sizeof(unsigned));
    for (int rep = 0; rep < 2; ++rep) k_probe<BODY><<<grid, 128>>>(outer, 0.999999, out, cyc, sm);
    cudaDeviceSynchronize();
    std::vector<long long> h(grid);
    std::vector<unsigned> hs(grid);
    cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hs.data(), sm, grid * sizeof(unsigned), cudaMemcpyDeviceToHost);
    std::vector<int> per_tpc(128, 0);
    for (unsigned s : hs) per_tpc[s / 2]++;
    int tpcs = 0, both = 0;
    for (int v : per_tpc) { tpcs += v > 0; both += v > 1; }
    double mean = 0, mean_pair = 0, mean_single = 0;
    int n_pair = 0, n_single = 0;
    for (int i = 0; i < grid; ++i) {
        mean += (double)h[i];
        if (per_tpc[hs[i] / 2] > 1) { mean_pair += (double)h[i]; ++n_pair; } else { mean_single += (double)h[i]; ++n_single; }
    }
    const long long mx = *std::max_element(h.begin(), h.end());
    const double per = (double)outer * BODY;
    printf("%-9s grid %3d: TPCs used %2d (both SMs busy in %2d)  cycles per DFMA (per warp): mean %.3f max %.3f | CTA alone in its TPC %.3f (n=%d), sharing it %.3f (n=%d)  [%s]\n",
           name, grid, tpcs, both, mean / grid / per, mx / per, n_single ? mean_single / n_single / per : 0.0, n_single,
           n_pair ? mean_pair / n_pair / per : 0.0, n_pair, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
    cudaFree(sm);
}

int main() {
    const long long total = 7168LL * 400;
    for (int grid : {1, 37, 74, 100, 148}) run<64>("loop", grid, total);
    for (int grid : {1, 37, 74, 100, 148}) run<7168>("straight", grid, total);
    for (int grid : {74, 148}) run<2048>("str32KB", grid, total);
    for (int grid : {74, 148}) run<1024>("str16KB", grid, total);
    return 0;
}
