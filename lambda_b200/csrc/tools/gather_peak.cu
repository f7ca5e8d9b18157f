// Microbenchmark: random-access read throughput of HBM3e on sm_100a at the granularities the FM-index
// seeding kernels use (seeding is bounded by random-sector throughput, not by streaming
// bandwidth; "report achieved rank-queries/s and sectors/s vs a random-gather microbenchmark").
//
// Every thread issues ILP independent random reads per iteration over a buffer far larger than L2
// (default 6 GiB, like the Li10 occurrence table of the 5M-sequence benchmark index):
//   sector32   one aligned 32-B sector                       (2 x 16 B)
//   line64     one aligned 64-B half line                    (4 x 16 B)
//   line128    one aligned 128-B line                        (8 x 16 B)
//   rank80     the reference layout: 80-B block at 80*b, read u32 count at +4*s and 32 B planes at +48
//   rank64     the device layout "occ64": 64-B aligned block, read u16 count at +32+2*s and 32 B planes at +0
// Output: one JSON object; G accesses/s and the useful GB/s of each pattern.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 64;
constexpr int ILP  = 8;

__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

template <int MODE>
__global__ void gather(unsigned char const * __restrict__ buf, unsigned long long nUnits, unsigned long long * out,
                       unsigned long long seed)
{
    unsigned long long const tid = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    unsigned long long       acc = 0;
    for (int it = 0; it < ITER; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
        {
            unsigned long long const h = mix(seed + tid * (ITER * ILP) + it * ILP + i);
            unsigned long long const u = h & (nUnits - 1); // nUnits is a power of two
            if (MODE == 0)
            {
                uint4 const * p = reinterpret_cast<uint4 const *>(buf + u * 32);
                uint4 const   a = __ldg(p), b = __ldg(p + 1);
                acc += a.x ^ a.w ^ b.y ^ b.w;
            }
            else if (MODE == 1)
            {
                uint4 const * p = reinterpret_cast<uint4 const *>(buf + u * 64);
                uint4 const   a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
                acc += a.x ^ b.y ^ c.z ^ d.w;
            }
            else if (MODE == 2)
            {
                uint4 const * p = reinterpret_cast<uint4 const *>(buf + u * 128);
                unsigned      x = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    x ^= __ldg(p + k).x;
                acc += x;
            }
            else if (MODE == 3)
            {
                unsigned char const *      b = buf + u * 80;
                unsigned int const         s = (static_cast<unsigned int>(h >> 40) & 7u) + 1u;
                unsigned int const         c = __ldg(reinterpret_cast<unsigned int const *>(b) + s);
                unsigned long long const * pl = reinterpret_cast<unsigned long long const *>(b + 48);
                acc += c + (__ldg(pl) ^ __ldg(pl + 1) ^ __ldg(pl + 2) ^ __ldg(pl + 3));
            }
            else
            {
                unsigned char const * b = buf + u * 64;
                unsigned int const    s = (static_cast<unsigned int>(h >> 40) & 7u) + 1u;
                uint4 const *         p = reinterpret_cast<uint4 const *>(b);
                uint4 const           a = __ldg(p), d = __ldg(p + 1);
                unsigned int const    c = __ldg(reinterpret_cast<unsigned short const *>(b + 32) + s);
                acc += c + (a.x ^ a.z ^ d.y ^ d.w);
            }
        }
    }
    if (acc == 0x1234567812345678ull)
        out[0] = acc;
}

template <int MODE>
static int run(char const * name, unsigned char const * buf, unsigned long long bytes, unsigned unit, unsigned useful,
               unsigned long long * out, int sms, bool last)
{
    unsigned long long nUnits = 1;
    while (nUnits * 2 * unit <= bytes)
        nUnits *= 2;
    int const blocks = sms * 64, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    gather<MODE><<<blocks, threads>>>(buf, nUnits, out, 1);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r)
    {
        cudaEventRecord(e0);
        gather<MODE><<<blocks, threads>>>(buf, nUnits, out, 1000 + r);
        cudaEventRecord(e1);
        CHECK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best)
            best = ms;
    }
    double const n = static_cast<double>(blocks) * threads * ITER * ILP;
    printf("\"%s\": {\"G_accesses_per_s\": %.2f, \"useful_GBps\": %.1f, \"unit_bytes\": %u, \"useful_bytes\": %u}%s\n", name,
           n / best / 1e6, n * useful / best / 1e6, unit, useful, last ? "" : ",");
    return 0;
}

int main(int argc, char ** argv)
{
    unsigned long long const bytes = (argc > 1 ? strtoull(argv[1], nullptr, 10) : 6ull) << 30;
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    unsigned char * buf;
    unsigned long long * out;
    CHECK(cudaMalloc(&buf, bytes));
    CHECK(cudaMemset(buf, 1, bytes));
    CHECK(cudaMalloc(&out, 8));
    int const sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"buffer_GiB\": %llu, \"threads_in_flight\": %d, \"ilp\": %d,\n", prop.name, sms,
           bytes >> 30, sms * 2048, ILP);
    if (run<0>("sector32", buf, bytes, 32, 32, out, sms, false)) return 1;
    if (run<1>("line64", buf, bytes, 64, 64, out, sms, false)) return 1;
    if (run<2>("line128", buf, bytes, 128, 128, out, sms, false)) return 1;
    if (run<3>("rank80", buf, bytes, 80, 36, out, sms, false)) return 1;
    if (run<4>("rank64", buf, bytes, 64, 34, out, sms, true)) return 1;
    printf("}\n");
    return 0;
}
