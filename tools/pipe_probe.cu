// Issue-rate probe (sm_100a): how many warp instructions per cycle one SM sub-partition sustains for a given mix, with
// plenty of warps and 8 independent chains per thread (no latency limit).  Answers whether the alu pipe (LOP3 / SHF /
// VIMNMX / SEL / PRMT), the fma pipe (IMAD) and the load/store unit overlap or share one issue budget.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MIX>
__global__ void probe(uint32_t seed, uint32_t mul, uint32_t* out, long long* cyc) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 7u;
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed + threadIdx.x * 8 + k;
    const uint32_t lane = threadIdx.x & 31;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MIX == 0) a[k] = (a[k] ^ seed) & (a[(k + 1) & 7] | 0x55u);                     // 1 LOP3
            if (MIX == 1) a[k] = a[k] * mul + a[(k + 1) & 7];                                  // 1 IMAD
            if (MIX == 2) { a[k] = (a[k] ^ seed) & (a[(k + 1) & 7] | 0x55u); a[k] = a[k] * mul + seed; }  // LOP3 + IMAD
            if (MIX == 3) a[k] = sm[(a[k] & 0xfe0u) | lane];                                   // LOP3 + LDS (conflict-free)
            if (MIX == 4) a[k] = min(a[k], a[(k + 1) & 7]) + 1u;                               // VIMNMX + IADD
            if (MIX == 5) a[k] = __byte_perm(a[k], a[(k + 1) & 7], 0x3215);                    // PRMT
            if (MIX == 6) a[k] = __funnelshift_r(a[k], a[(k + 1) & 7], 7);                     // SHF
            if (MIX == 7) a[k] = a[k] > seed ? a[(k + 1) & 7] : a[k] + 3u;                     // ISETP + SEL (+IADD)
            if (MIX == 9) a[k] = (uint32_t)__double2hiint(__ull2double_rz(((uint64_t)a[k] << 20) | a[(k + 1) & 7])) + seed;  // I2F.F64.U64 (+IADD)
            if (MIX == 10) a[k] = (uint32_t)__double2hiint((double)a[k]) ^ a[(k + 1) & 7];       // I2F.F64.U32 (+LOP3)
            if (MIX == 11) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(__hiloint2double((int)(a[k] | 0x40000000u), 0))); a[k] = (uint32_t)__double2hiint(r) + seed; }  // MUFU.RCP64H (+IADD)
            if (MIX == 12) { double r = fma(__hiloint2double((int)(a[k] & 0x400fffffu) | 0x3ff00000, (int)seed), 1.0000001, 0.5); a[k] = (uint32_t)__double2loint(r) ^ (uint32_t)__double2hiint(r); }  // DFMA (+2 LOP3)
            if (MIX == 13) { const uint64_t w = (uint64_t)a[k] * mul + a[(k + 1) & 7]; a[k] = (uint32_t)w ^ (uint32_t)(w >> 32); }  // IMAD.WIDE (+LOP3)
            if (MIX == 14) { const uint64_t w = __umul64hi(((uint64_t)a[k] << 32) | seed, ((uint64_t)mul << 32) | a[(k + 1) & 7]); a[k] = (uint32_t)w ^ (uint32_t)(w >> 32); }  // mul.hi.u64 (+LOP3)
            if (MIX == 8) { a[k] = (a[k] ^ seed) & (a[(k + 1) & 7] | 0x55u); a[k] = sm[(a[k] & 0xfe0u) | lane]; a[k] = a[k] * mul + seed; }  // LOP3,LOP3,LDS,IMAD
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MIX] = t1 - t0;
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 32 * sizeof(long long));
    const char* names[] = {"LOP3", "IMAD", "LOP3+IMAD", "LOP3+LDS", "VIMNMX+IADD", "PRMT", "SHF", "ISETP+SEL(+IADD)", "LOP3,LOP3,LDS,IMAD", "I2F.F64.U64+IADD", "I2F.F64.U32+LOP3", "MUFU.RCP64H+IADD", "DFMA+2LOP3", "IMAD.WIDE+LOP3", "mul.hi.u64+LOP3"};
    for (int warps = 4; warps <= 32; warps *= 2) {  // warps per SM (one CTA per SM)
#define RUN(K) probe<K><<<148, warps * 32>>>(12345u, 3u, out, cyc);
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14)
        cudaDeviceSynchronize();
        for (int k = 0; k < 15; k++)
            printf("warps/SM %2d  %-22s %9.2f cycles per loop trip\n", warps, names[k], (double)cyc[k] / ITERS);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
