// Microbenchmark: can packed-fp16 arithmetic on the FMA pipe run next to the packed-int16 DPX
// instructions on the ALU pipe?  (Idea: a second formulation of the DP cell update in fp16x2 --
// max(a,b) = a + relu(b - a), exact for |values| <= 2048 -- for warps that share the SM with the DPX
// warps.)  Reports warp-instructions/s for:
//   hfma2, hfma2_relu, hadd2          alone (every warp)
//   hmnmx2                            alone (does half2 max live on the ALU pipe?)
//   dpx_only                          VIADDMNMX.S16x2 in every warp
//   mix_warps                         warps 0-3 of a block VIADDMNMX, warps 4-7 HFMA2.RELU: total rate
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 4096;
constexpr int ILP  = 8;

__device__ __forceinline__ unsigned hfma2_relu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned hfma2(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned hadd2(unsigned a, unsigned b)
{
    unsigned d;
    asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned hmax2(unsigned a, unsigned b)
{
    unsigned d;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// OP: 0 hfma2, 1 hfma2.relu, 2 hadd2, 3 hmax2, 4 viaddmnmx, 5 mix by warp parity (even: viaddmnmx, odd: hfma2.relu)
template <int OP>
__global__ void kern(unsigned * out, unsigned a, unsigned b)
{
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        x[i] = threadIdx.x * 7u + i + a;
    bool const odd = (threadIdx.x >> 7) & 1; // warps 0-3 vs 4-7 of a block: every scheduler (warp id % 4) gets both kinds
    if ((OP == 5 || OP == 8) && odd)
    {
        for (int it = 0; it < ITER; ++it)
#pragma unroll
            for (int i = 0; i < ILP; ++i)
                x[i] = (OP == 5) ? hfma2_relu(x[i], a, b) : x[i] * a + b;
    }
    else
    {
        for (int it = 0; it < ITER; ++it)
        {
#pragma unroll
            for (int i = 0; i < ILP; ++i)
            {
                if (OP == 0) x[i] = hfma2(x[i], a, b);
                else if (OP == 1) x[i] = hfma2_relu(x[i], a, b);
                else if (OP == 2) x[i] = hadd2(x[i], a);
                else if (OP == 3) x[i] = hmax2(x[i], b);
                else if (OP == 6) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = hfma2_relu(x[i], a, b); } // same warp, both pipes
                else if (OP == 7) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = x[i] * a + b; }          // DPX + IMAD
                else x[i] = __viaddmax_s16x2(x[i], a, b);
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        s ^= x[i];
    if (s == 0x12345678u)
        out[0] = s;
}

template <int OP>
static double run(unsigned * d, int sms)
{
    int const blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<OP><<<blocks, threads>>>(d, 0x3c003c00u, 0x40004000u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r)
    {
        cudaEventRecord(e0);
        kern<OP><<<blocks, threads>>>(d, 0x3c003c00u, 0x40004000u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return double(blocks) * (threads / 32) * double(ITER) * ILP / (best * 1e-3);
}

int main()
{
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    unsigned * d;
    CHECK(cudaMalloc(&d, 4));
    int const sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"unit\": \"G warp-instr/s\", \"hfma2\": %.1f, \"hfma2_relu\": %.1f, \"hadd2\": %.1f, \"hmax2\": %.1f, "
           "\"viaddmnmx\": %.1f, \"warps_viaddmnmx_and_warps_hfma2relu_total\": %.1f, \"same_warp_viaddmnmx_hfma2relu_total\": %.1f, "
           "\"same_warp_viaddmnmx_imad_total\": %.1f, \"warps_viaddmnmx_and_warps_imad_total\": %.1f}\n",
           p.name, run<0>(d, sms) / 1e9, run<1>(d, sms) / 1e9, run<2>(d, sms) / 1e9, run<3>(d, sms) / 1e9, run<4>(d, sms) / 1e9,
           run<5>(d, sms) / 1e9, 2 * run<6>(d, sms) / 1e9, 2 * run<7>(d, sms) / 1e9, run<8>(d, sms) / 1e9);
    return 0;
}
