// Microbenchmark: sustained throughput of the packed-int16 instruction classes the extension DP
// uses on sm_100a (MEASURED_PEAKS.json has no integer peak, the builder commits one).
//
// Each test runs a long dependent-free stream of one SASS instruction class from every resident
// warp and reports warp-instructions/s for the whole GPU.  Output: one JSON object on stdout.
//   VIADDMNMX.S16x2[.RELU]  (__viaddmax_s16x2[_relu])   add + max on two int16 lanes
//   VIMNMX3.S16x2           (__vimax3_s16x2)
//   VIMNMX.S16x2            (__vmaxs2)
//   VIADD.16x2              (__vadd2)
//   PRMT                    (__byte_perm)
//   IMAD / IADD3 32-bit     for the pipe split
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 4096;
constexpr int ILP  = 8;

template <int OP>
__global__ void kern(unsigned * out, unsigned a, unsigned b)
{
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        x[i] = threadIdx.x * 7u + i + a;
    for (int it = 0; it < ITER; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
        {
            if (OP == 0) x[i] = __viaddmax_s16x2(x[i], a, b);
            else if (OP == 1) x[i] = __viaddmax_s16x2_relu(x[i], a, b);
            else if (OP == 2) x[i] = __vimax3_s16x2(x[i], a, b);
            else if (OP == 3) x[i] = __vmaxs2(x[i], b);
            else if (OP == 4) x[i] = __vadd2(x[i], a);
            else if (OP == 5) x[i] = __byte_perm(x[i], a, 0x5140 ^ (b & 0x1111));
            else if (OP == 6) x[i] = x[i] * a + b;          // IMAD
            else if (OP == 7) x[i] = (x[i] + a) ^ b;         // IADD3 + LOP3 -> 2 instr
            else if (OP == 8) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = x[i] * a + b; } // ALU + FMA pipes together
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
static double run(unsigned * d, int sms, int instrPerIter)
{
    int const blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<OP><<<blocks, threads>>>(d, 3, 5);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r)
    {
        cudaEventRecord(e0);
        kern<OP><<<blocks, threads>>>(d, 3, 5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double const warpInstr = double(blocks) * (threads / 32) * double(ITER) * ILP * instrPerIter;
    return warpInstr / (best * 1e-3); // warp-instructions per second
}

int main()
{
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    unsigned * d;
    CHECK(cudaMalloc(&d, 4));
    int const sms = p.multiProcessorCount;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    char const * names[] = {"viaddmnmx_s16x2", "viaddmnmx_s16x2_relu", "vimnmx3_s16x2", "vimnmx_s16x2", "viadd_16x2",
                            "prmt", "imad", "iadd3_lop3_pair", "viaddmnmx_plus_imad"};
    double r[9];
    r[0] = run<0>(d, sms, 1);
    r[1] = run<1>(d, sms, 1);
    r[2] = run<2>(d, sms, 1);
    r[3] = run<3>(d, sms, 1);
    r[4] = run<4>(d, sms, 1);
    r[5] = run<5>(d, sms, 1);
    r[6] = run<6>(d, sms, 1);
    r[7] = run<7>(d, sms, 2);
    r[8] = run<8>(d, sms, 2);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_khz\": %d, \"unit\": \"G warp-instr/s (x32 lanes, x2 int16 halves)\"", p.name, sms, clk);
    for (int i = 0; i < 9; ++i)
        printf(", \"%s\": %.2f", names[i], r[i] / 1e9);
    // lane-instr per clk per SM at the max clock, for orientation
    printf(", \"viaddmnmx_lanes_per_clk_per_sm_at_max_clock\": %.2f", r[0] * 32 / (double(clk) * 1e3) / sms);
    printf("}\n");
    return 0;
}
