// tools/fp64_probe.cu — microbenchmark: FP64 add/mul issue rate and dependent-issue latency on one SM.
// Not part of the product; used once to size the blocked kernel (DESIGN.md "FP64 ceiling").
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_chain(double* out, int iters, double a, double b, long long* cyc) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __dadd_rn(__dmul_rn(v[i], b), a);  // 2 dependent FP64 ops, no FMA
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int nsm) {
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024 * 1024);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm);
    const int blocks_per_sm = (warps_per_sm * 32 + threads - 1) / threads;
    k_chain<ILP><<<nsm * blocks_per_sm, threads>>>(out, 16, 1.0, 1.0000001, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_chain<ILP><<<nsm * blocks_per_sm, threads>>>(out, iters, 1.0, 1.0000001, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops_per_warp = 2.0 * ILP * iters;                 // warp-level FP64 instructions
    const double per_sm_per_cycle = ops_per_warp * warps_per_sm / (double)c;
    printf("ILP %d warps/SM %2d: %lld cycles, %.3f FP64 warp-instr/cycle/SM (%.1f lanes/clk/SM), "
           "%.2f cycles per dependent op per warp, chip %.2f T lane-ops/s\n",
           ILP, warps_per_sm, c, per_sm_per_cycle, per_sm_per_cycle * 32, (double)c / (2.0 * iters) ,
           ops_per_warp * 32 * warps_per_sm * blocks_per_sm / blocks_per_sm * nsm / (ms * 1e-3) / 1e12);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {1, 4, 8, 16, 32}) run<1>(w, p.multiProcessorCount);
    for (int w : {4, 8, 16, 32}) run<2>(w, p.multiProcessorCount);
    for (int w : {4, 8, 16, 32}) run<4>(w, p.multiProcessorCount);
    for (int w : {4, 8, 16}) run<8>(w, p.multiProcessorCount);
    return 0;
}
