// Dependent-chain latency probe for the instructions on the rANS critical path (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void probe(uint64_t seed, uint32_t f, double inv, uint64_t* out, long long* cyc) {
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = (i * 7 + 1) & 2047;
    __syncthreads();
    uint64_t x = seed + threadIdx.x;
    uint32_t a = (uint32_t)seed | 1u, b = f;
    double d = (double)seed;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __double2ull_rz(__ull2double_rz(x) * inv) + (x & 0xfff0000000000000ull);          // I2F64+DMUL+F2I64 (+LOP/IADD)
        if (OP == 1) d = __ull2double_rz(__double_as_longlong(d));                                           // I2F.F64.U64 only
        if (OP == 2) x = __double2ull_rz(__longlong_as_double(x | 0x4000000000000000ull));                   // F2I.U64.F64 only
        if (OP == 3) d = d * inv;                                                                            // DMUL
        if (OP == 4) a = __umulhi(a, b) + a;                                                                 // IMAD.HI
        if (OP == 5) x = (uint64_t)(uint32_t)x * b + x;                                                      // IMAD.WIDE
        if (OP == 6) a = sm[a & 2047];                                                                       // LDS
        if (OP == 7) a = (a >> 3) ^ b;                                                                       // SHF+LOP
        if (OP == 8) a = a * b + 7;                                                                          // IMAD
        if (OP == 9) { float r = __uint_as_float(a); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(r)); a = __float_as_uint(r) | 0x3f000000u; }  // MUFU.RCP
        if (OP == 10) a = __float2uint_rz(__uint2float_rn(a) * 0.999f);                                      // I2F.F32 + FMUL + F2I.U32
        if (OP == 11) a = (a >= b) ? a - b : a + 3;                                                          // ISETP+SEL
        if (OP == 12) x = (x << 15) | (x >> 49);                                                             // 64-bit shift pair
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + a + (uint64_t)d;
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}
int main() {
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 16 * sizeof(long long));
    const char* names[] = {"I2F64.U64+DMUL+F2I.U64.F64(+2 alu)", "I2F.F64.U64", "F2I.U64.F64", "DMUL", "IMAD.HI(+IADD)", "IMAD.WIDE", "LDS(+LOP)", "SHF+LOP", "IMAD",
                           "MUFU.RCP(+LOP)", "I2F.F32+FMUL+F2I.U32", "ISETP+SEL(+IADD)", "SHF64 pair"};
#define RUN(K) probe<K><<<1, 32>>>(0x123456789abcdefull, 12345u, 1.0 / 12345.0, out, cyc);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    cudaDeviceSynchronize();
    for (int k = 0; k < 13; k++) printf("%-40s %.1f cycles/iter\n", names[k], (double)cyc[k] / N);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
