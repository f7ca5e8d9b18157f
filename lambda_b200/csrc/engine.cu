// lambda_b200 engine: device-resident index, per-context pipeline, C ABI (include/lambda_b200.h).
//
// One lgpu_search_batch() call = the reference's batch loop body (src/search.cpp:428-457) for an
// arbitrarily large batch:
//
//   H2D queries -> prepQueries (frames, reduction)
//   phase 1 (searchOpts0) on all queries, phase 2 (searchOpts) on the queries without a hit:
//     seedKernel            search()                         src/search_algo.hpp:607-762
//     widen/sort/merge      _widenAndPreprocessMatches       :1137-1175
//     swWavefront<score>    _performAlignment<false>         :1246
//     filterKernel          bit-score / e-value thresholds   :1252-1281 (integer thresholds from host)
//     swWavefront<trace>    _performAlignment<true>          :1296
//     tracebackKernel       _expandAlign + computeAlignmentStats :1306-1308
//   D2H hits -> host: bit score / e-value doubles, identity cut-off, phase bookkeeping
//   (iterativeSearchPre/Post :1391-1460), optional _writeRecord finalisation (:821-913).
//
// No CPU fallback exists: if CUDA is unavailable every entry point fails with LGPU_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include "../../include/lambda_b200.h"
#include "host_finalize.hpp"
#include "host_params.hpp"
#include "kernels_dpx.cuh"
#include "kernels_dpx_trace.cuh"
#include "kernels_extend.cuh"
#include "kernels_finalize.cuh"
#include "kernels_fm.cuh"
#include "lba_index.hpp"

namespace lgpu
{

struct CudaError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};
struct UnsupportedError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

#define LGPU_CUDA(expr)                                                                                               \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t const e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                        \
            throw ::lgpu::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_));                              \
    } while (0)

static thread_local std::string g_lastError;

// LAMBDA_B200_TRACE_TIMES=1: wall-clock marks on stderr (where does a cold first call spend its time?)
static void traceTime(char const * label)
{
    static int const on = [] {
        char const * e = std::getenv("LAMBDA_B200_TRACE_TIMES");
        return e && std::atoi(e) != 0 ? 1 : 0;
    }();
    if (!on)
        return;
    static auto const t0 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[lgpu %9.3f ms] %s\n",
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), label);
}

// growable device buffer
template <typename T>
struct DevBuf
{
    T *    p   = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    DevBuf()                           = default;
    DevBuf(DevBuf const &)             = delete;
    DevBuf & operator=(DevBuf const &) = delete;
    void     release()
    {
        if (p)
            cudaFree(p);
        p   = nullptr;
        cap = 0;
    }
    void reserve(size_t n)
    {
        if (n <= cap)
            return;
        release();
        size_t const want = std::max<size_t>(n + n / 4, 256);
        LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T)));
        cap = want;
    }
    // like reserve(), but the first `keep` elements survive a reallocation
    void grow(size_t n, size_t keep, cudaStream_t s)
    {
        if (n <= cap)
            return;
        T *          old  = p;
        size_t const want = std::max<size_t>(n + n / 2, 256);
        T *          np   = nullptr;
        LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&np), want * sizeof(T)));
        if (old && keep)
        {
            LGPU_CUDA(cudaMemcpyAsync(np, old, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
            LGPU_CUDA(cudaStreamSynchronize(s));
        }
        if (old)
            cudaFree(old);
        p   = np;
        cap = want;
    }
};

// growable pinned host staging buffer (async copies from/to pageable memory are staged and slow)
template <typename T>
struct PinnedBuf
{
    T *    p   = nullptr;
    size_t cap = 0;
    ~PinnedBuf()
    {
        if (p)
            cudaFreeHost(p);
    }
    PinnedBuf()                              = default;
    PinnedBuf(PinnedBuf const &)             = delete;
    PinnedBuf & operator=(PinnedBuf const &) = delete;
    void        reserve(size_t n)
    {
        if (n <= cap)
            return;
        if (p)
            cudaFreeHost(p);
        p                 = nullptr;
        cap               = 0;
        size_t const want = std::max<size_t>(n + n / 4, 256);
        LGPU_CUDA(cudaMallocHost(reinterpret_cast<void **>(&p), want * sizeof(T)));
        cap = want;
    }
};

struct HostTimer
{
    float *                               acc;
    std::chrono::steady_clock::time_point t0;
    explicit HostTimer(float * a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer()
    {
        if (acc)
            *acc += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

template <typename T>
static T * uploadNew(T const * host, size_t n)
{
    T * d = nullptr;
    LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&d), std::max<size_t>(n, 1) * sizeof(T)));
    if (n)
        LGPU_CUDA(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

// Host -> device copies of the index blobs.  The source is the mmap'd .lba (pageable page cache): a plain
// cudaMemcpy moves it through the driver's small staging buffer at a few GB/s.  Here several host
// threads copy 16 MiB chunks into their own pinned double buffers and push them with cudaMemcpyAsync, so
// the page-cache reads of all threads and the PCIe transfers overlap.
struct UploadJob
{
    unsigned char *       dst;
    unsigned char const * src;
    size_t                bytes;
};

static void runUploads(std::vector<UploadJob> const & jobs, int device)
{
    size_t kChunk = 16u << 20;
    if (char const * e = std::getenv("LAMBDA_B200_UPLOAD_CHUNK_MB"))
        kChunk = static_cast<size_t>(std::max(1, std::min(256, std::atoi(e)))) << 20;
    struct Piece
    {
        unsigned char *       dst;
        unsigned char const * src;
        size_t                bytes;
        int                   fd;      // >= 0: the source is a mapped index file, read it with pread()
        size_t                fileOff;
    };
    std::vector<Piece> pieces;
    for (UploadJob const & j : jobs)
    {
        int    fd      = -1;
        size_t fileOff = 0;
        if (!lbaLocate(j.src, j.bytes, fd, fileOff))
            fd = -1;
        for (size_t off = 0; off < j.bytes; off += kChunk)
            pieces.push_back({j.dst + off, j.src + off, std::min(kChunk, j.bytes - off), fd, fileOff + off});
    }
    if (pieces.empty())
        return;
    // page-cache -> pinned copies are what limits the upload: one thread moves ~8 GB/s, PCIe Gen5 takes ~50 GB/s
    unsigned int nThreads = std::max(4u, std::min(12u, std::thread::hardware_concurrency() * 3 / 4));
    if (char const * e = std::getenv("LAMBDA_B200_UPLOAD_THREADS"))
        nThreads = static_cast<unsigned int>(std::max(1, std::min(32, std::atoi(e))));
    nThreads = static_cast<unsigned int>(std::min<size_t>(nThreads, pieces.size()));
    char const * const um           = std::getenv("LAMBDA_B200_UPLOAD_MODE");
    bool const         registerMode = um && !std::strcmp(um, "register");
    std::atomic<size_t>      next{0};
    std::vector<std::string> err(nThreads);
    std::vector<std::thread> th;
    for (unsigned int t = 0; t < nThreads; ++t)
        th.emplace_back([&, t] {
            try
            {
                LGPU_CUDA(cudaSetDevice(device));
                cudaStream_t    st;
                cudaEvent_t     ev[2];
                unsigned char * pin[2] = {nullptr, nullptr};
                LGPU_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                for (int b = 0; b < 2; ++b)
                {
                    LGPU_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&pin[b]), kChunk, cudaHostAllocDefault));
                    LGPU_CUDA(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
                }
                bool used[2] = {false, false};
                int  b       = 0;
                for (;;)
                {
                    size_t const i = next.fetch_add(1);
                    if (i >= pieces.size())
                        break;
                    if (used[b])
                        LGPU_CUDA(cudaEventSynchronize(ev[b]));
                    Piece const & pc   = pieces[i];
                    bool          done = false;
                    if (registerMode && pc.fd >= 0)
                    {
                        // experiment (LAMBDA_B200_UPLOAD_MODE=register): pin the mapped pages themselves and let the copy
                        // engine read the page cache directly -- no host-side copy at all
                        void * const src = const_cast<unsigned char *>(pc.src);
                        if (cudaHostRegister(src, pc.bytes, cudaHostRegisterReadOnly) == cudaSuccess)
                        {
                            cudaError_t const e1 = cudaMemcpyAsync(pc.dst, src, pc.bytes, cudaMemcpyHostToDevice, st);
                            cudaError_t const e2 = cudaStreamSynchronize(st);
                            cudaHostUnregister(src);
                            LGPU_CUDA(e1);
                            LGPU_CUDA(e2);
                            continue;
                        }
                        cudaGetLastError();
                    }
                    if (pc.fd >= 0)
                    {
                        size_t got = 0;
                        while (got < pc.bytes)
                        {
                            ssize_t const r = pread(pc.fd, pin[b] + got, pc.bytes - got, static_cast<off_t>(pc.fileOff + got));
                            if (r <= 0)
                                break;
                            got += static_cast<size_t>(r);
                        }
                        done = got == pc.bytes;
                    }
                    if (!done)
                        std::memcpy(pin[b], pc.src, pc.bytes);
                    LGPU_CUDA(cudaMemcpyAsync(pieces[i].dst, pin[b], pieces[i].bytes, cudaMemcpyHostToDevice, st));
                    LGPU_CUDA(cudaEventRecord(ev[b], st));
                    used[b] = true;
                    b ^= 1;
                }
                LGPU_CUDA(cudaStreamSynchronize(st));
                for (int k = 0; k < 2; ++k)
                {
                    cudaFreeHost(pin[k]);
                    cudaEventDestroy(ev[k]);
                }
                cudaStreamDestroy(st);
            }
            catch (std::exception const & e)
            {
                err[t] = e.what();
            }
        });
    for (auto & t : th)
        t.join();
    for (auto const & e : err)
        if (!e.empty())
            throw CudaError("index upload: " + e);
}

static inline unsigned int gridFor(unsigned long long n, unsigned int block)
{
    return static_cast<unsigned int>((n + block - 1) / block);
}

} // namespace lgpu

using namespace lgpu;

// -------------------------------------------------------------------------------------------------
// index
// -------------------------------------------------------------------------------------------------

struct lgpu_lba
{
    std::unique_ptr<LbaFile> file;
};

struct lgpu_index
{
    int                      device = 0;
    DevIndex                 dev{};
    lgpu_index_desc          meta{}; // scalar fields only (host pointers cleared)
    std::vector<void *>      allocs;
    std::vector<size_t>      allocBytes; // size of every entry of `allocs` (lgpu_index_clone)
    uint64_t                 bytes         = 0;
    uint64_t                 dbTotalLength = 0;
    std::vector<uint64_t>    sbjDelimsHost; // host copy of dev.seqDelims (window checks of the stage API)
    // depth-k prefix tables of the seeding kernels (kernels_fm.cuh seedPrefixCursor), built on first use
    mutable std::mutex                      prefixMutex;
    mutable std::map<unsigned int, uint2 *> prefixTabs;
    ~lgpu_index()
    {
        cudaSetDevice(device);
        for (void * p : allocs)
            cudaFree(p);
        for (auto & kv : prefixTabs)
            cudaFree(kv.second);
    }
};

struct lgpu_ctx
{
    lgpu_index const * index = nullptr;
    int                device = 0;
    lgpu_params        params{};
    Scoring            scoring;
    DomainInfo         di;
    cudaStream_t       stream = nullptr;
    std::string        err;
    int                numSMs = 148;

    // device buffers
    DevBuf<signed char>        dMatrix;
    DevBuf<unsigned char>      dQOrig, dQTrans, dQRed;
    DevBuf<unsigned long long> dQOffs;
    DevBuf<unsigned int>       dActive;
    DevBuf<lgpu_match>         dMatches, dMerged, dTasks2, dUserMatches;
    DevBuf<unsigned long long> dCounters;
    DevBuf<unsigned long long> dKey1, dKey2, dKey1b, dKey2b;
    DevBuf<unsigned int>       dPerm, dPermB, dHead, dScan;
    DevBuf<unsigned char>      dCubTemp;
    DevBuf<int>                dScores, dScores2, dMinBit, dMinEval;
    DevBuf<unsigned int>       dWork, dBestPos, dBoundary, dOrder, dOrderB, dClassInfo, dSegStart, dSegStartB, dJobHead, dJobPos, dJobs;
    DevBuf<unsigned long long> dClassKeys, dClassKeysB;
    bool                       dpxOk = false; // scoring fits the int8 profile of the DPX kernels
    unsigned int               dpxBlocksPerSM = 32; // resident warps of the DP score kernel per SM (LAMBDA_B200_DPX_OCC)
    bool                       dpxScoreOk = false; // ... and every matrix entry >= gap open (score kernel: profile bytes >= 0)
    uint64_t                   minSubBatch = 16384; // queries per sub-batch below which a call is not cut further
    unsigned int               streams = 1;  // sub-batches in flight per lgpu_search_batch call (LAMBDA_B200_STREAMS)
    std::vector<std::unique_ptr<lgpu_ctx>> workers;
    int                        seedMode = 0; // LAMBDA_B200_SEED=thread|warp|block|spec forces one seeding kernel (tests); 0 = auto
    bool                       seedTextElong = true; // unique cursors: locate once, compare texts (LAMBDA_B200_SEED_TEXT=0: LF steps)
    bool                       seedPrefix    = true; // prefix table in front of the exact part of a seed (LAMBDA_B200_SEED_PREFIX=0: off)
    DevBuf<unsigned long long> dSeedCursors;
    DevBuf<unsigned int>       dSeedCounts;
    DevBuf<unsigned char>      dTrace;
    DevBuf<unsigned long long> dTraceOff;
    DevBuf<lgpu_hit>           dHits;
    DevBuf<unsigned int>       dPlanes;      // residue planes of DP pass 2 (one byte per cell)
    DevBuf<lgpu_match>         dTasksScalar;
    bool                       forceScalarTrace = false; // LAMBDA_B200_TRACE=scalar (tests)
    bool                       privProfiles = false; // DP pass 1: one query profile per group (small alphabets) instead of per warp
    bool                       resTraceOk   = false; // scoring fits the residue-plane trace path (kernels_dpx_trace.cuh)
    int                        traceTab     = kDpxTabTrace32; // class table of DP pass 2
    uint64_t                   maxPlaneWords = (16ull << 30) / 4; // residue planes per launch group (LAMBDA_B200_PLANE_MB)
    unsigned int               maxColsInt16 = 0; // longest query whose scores cannot leave int16 (32767 / largest matrix entry)
    DevBuf<unsigned int>       dOverflow;        // [0] != 0: an alignment score exceeded 32767 (the reference wraps there)
    PinnedBuf<unsigned int>    hOverflow;
    DevBuf<unsigned long long> dPlaneWords, dPlaneWordsB;
    DevBuf<unsigned int>       dScalarSlots, dScalarIdx;

    // records of the batch on the device: all phases appended, then finalised (kernels_finalize.cuh)
    DevBuf<lgpu_hit>           dAllHits, dFinal;
    DevBuf<unsigned int>       dQryHasHit, dFinIdx, dFinIdxB, dFinIdx1;
    DevBuf<unsigned char>      dDropped;
    uint64_t                   nAll = 0, nFinal = 0;
    lgpu_hit const *           finalDev = nullptr; // the nFinal records of the last searchOne on the device
    unsigned int               maxQueryLen = 0;
    std::unique_ptr<EValueComputer> evc;           // persistent: length adjustments and score tables are cached
    std::vector<ScoreThresholds>    thrByLen;      // thresholds per query length (direct index), filled lazily
    std::vector<unsigned char>      thrKnown;
    std::unordered_map<uint64_t, ScoreThresholds> thrLong; // ... for lengths beyond the table
    PinnedBuf<int>                  hMinBit, hMinEval;
    struct ExportPart
    {
        lgpu_ctx const * ctx;
        uint64_t         qBase, cigarBase, n;
    };
    std::vector<ExportPart>    lastParts; // where the records of the last lgpu_search_batch live (lgpu_ctx_export_hits)

    // side streams for the per-class launches of the DP passes (classStream / joinClassStreams)
    static constexpr int kAux = 3;
    cudaStream_t         aux[kAux]{};
    cudaEvent_t          auxDone[kAux]{};
    unsigned int         auxUsed      = 0;
    bool                 classStreams = true; // LAMBDA_B200_CLASS_STREAMS=0: every class on the context's stream

    // stage timers: event pairs recorded on the stream, read once at the end of the call
    struct TimerSlot
    {
        cudaEvent_t a = nullptr, b = nullptr;
        float *     acc = nullptr;
    };
    std::vector<TimerSlot> timers;
    size_t                 timersUsed = 0;

    // pinned staging
    PinnedBuf<lgpu_match> hTasks;
    PinnedBuf<lgpu_hit>   hHits;
    PinnedBuf<int>        hThresh;

    // host results: the records of the last call (pinned; sub-batch workers write their slices directly)
    PinnedBuf<lgpu_hit>     hResult;
    uint64_t                nResult = 0;
    std::vector<uint32_t>   cigar;       // run-length ops of the hits (lgpu_params.want_cigar), see lgpu_hit::cigar_off
    DevBuf<unsigned int>    dCigarCap, dCigarOff, dCigar;
    PinnedBuf<unsigned int> hCigarStage;
    std::vector<lgpu_match> matchesHost;
    cudaEvent_t             ev[8]{};
    cudaEvent_t             evSync = nullptr; // blocking-sync event: waiting host threads sleep instead of spinning
    int                     syncMode = 0;     // LGPU_SYNC_*: how host threads wait for the stream (LAMBDA_B200_SYNC)

    DevQueries Q{};
    uint64_t   nQueries = 0, totalResidues = 0;
    std::vector<uint64_t> qOffsHost;

    ~lgpu_ctx()
    {
        if (index)
            cudaSetDevice(device); // not index->device: the index may already be gone (e.g. garbage-collected first)
        for (auto & e : ev)
            if (e)
                cudaEventDestroy(e);
        for (int i = 0; i < kAux; ++i)
        {
            if (auxDone[i])
                cudaEventDestroy(auxDone[i]);
            if (aux[i])
                cudaStreamDestroy(aux[i]);
        }
        for (auto & t : timers)
        {
            if (t.a)
                cudaEventDestroy(t.a);
            if (t.b)
                cudaEventDestroy(t.b);
        }
        if (evSync)
            cudaEventDestroy(evSync);
        if (stream)
            cudaStreamDestroy(stream);
    }
};

namespace lgpu
{

// -------------------------------------------------------------------------------------------------
// pipeline stages
// -------------------------------------------------------------------------------------------------

// Host waits.  A step has ~50 points where a host thread waits for its stream, and with one process per GPU and
// three sub-batch threads per process a box may run more waiting threads than it has cores.
//   spin   cudaStreamSynchronize / cudaEventSynchronize (the driver spins): lowest latency, the measured best
//          while every waiting thread has a core of its own (58.7 ms per step on 1 GPU, 61.8 ms on 4 GPUs / 16 cores)
//   block  sleep on a blocking-sync event: measured slower (wake-up latency: 61.0 / 66.6 ms in the same runs)
//   yield  poll cudaStreamQuery / cudaEventQuery and sched_yield() between polls: behaves like spinning when the
//          core is not contended (sched_yield returns at once) and hands the core to a runnable thread -- e.g. a
//          sub-batch thread that has kernels to launch -- when it is
//   auto   (default) spin, unless the search threads of all ranks on this box outnumber the cores this process may
//          run on (LOCAL_WORLD_SIZE x threads inside searchOne > sched_getaffinity count): then yield
// LAMBDA_B200_SYNC=spin|block|yield|auto; LAMBDA_B200_BLOCKING_SYNC=1 is the older spelling of block.
// Measured (profiles/r1_sweep_hostwait.jsonl, searchp step, 1 GPU): all 16 cores: spin 59.0 ms, yield 59.3 ms;
// process pinned to 2 cores (3 search threads): spin 63.9 ms, auto (= yield) 64.1 ms, block 66.2 ms.
enum { LGPU_SYNC_AUTO = 0, LGPU_SYNC_SPIN = 1, LGPU_SYNC_BLOCK = 2, LGPU_SYNC_YIELD = 3 };
static std::atomic<int> g_searchThreads{0}; // host threads of this process currently inside searchOne

struct SearchThreadScope
{
    SearchThreadScope() { g_searchThreads.fetch_add(1, std::memory_order_relaxed); }
    ~SearchThreadScope() { g_searchThreads.fetch_sub(1, std::memory_order_relaxed); }
};

static int hostCores()
{
    static int const n = [] {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0)
            return CPU_COUNT(&set);
        unsigned int const h = std::thread::hardware_concurrency();
        return h ? static_cast<int>(h) : 1;
    }();
    return n;
}

static int localWorldSize()
{
    static int const n = [] {
        char const * e = std::getenv("LOCAL_WORLD_SIZE"); // set by torchrun: ranks on this box
        int const    v = e ? std::atoi(e) : 1;
        return v > 0 ? v : 1;
    }();
    return n;
}

static inline int effectiveSyncMode(lgpu_ctx const & c)
{
    if (c.syncMode != LGPU_SYNC_AUTO)
        return c.syncMode;
    int const waiting = std::max(1, g_searchThreads.load(std::memory_order_relaxed)) * localWorldSize();
    return waiting > hostCores() ? LGPU_SYNC_YIELD : LGPU_SYNC_SPIN;
}

template <typename Query>
static inline void pollYield(Query && query)
{
    for (unsigned int polls = 0;; ++polls)
    {
        cudaError_t const e = query();
        if (e == cudaSuccess)
            return;
        if (e != cudaErrorNotReady)
            LGPU_CUDA(e);
        if (polls >= 16)
            sched_yield();
    }
}

static inline void syncStream(lgpu_ctx & c)
{
    switch (effectiveSyncMode(c))
    {
        case LGPU_SYNC_BLOCK:
            LGPU_CUDA(cudaEventRecord(c.evSync, c.stream));
            LGPU_CUDA(cudaEventSynchronize(c.evSync));
            break;
        case LGPU_SYNC_YIELD:
            pollYield([&] { return cudaStreamQuery(c.stream); });
            break;
        default:
            LGPU_CUDA(cudaStreamSynchronize(c.stream));
    }
}

static inline void waitEvent(lgpu_ctx & c, cudaEvent_t ev)
{
    if (effectiveSyncMode(c) == LGPU_SYNC_YIELD)
        pollYield([&] { return cudaEventQuery(ev); });
    else
        LGPU_CUDA(cudaEventSynchronize(ev)); // block mode: the events were created with cudaEventBlockingSync
}

// Stage times: an event pair per stage instance, recorded on the stream and read once at the end of the call
// (resolveTimers) -- a timer never makes the host wait.
struct StageTimer
{
    lgpu_ctx & c;
    int        slot = -1;
    StageTimer(lgpu_ctx & ctx, float * acc) : c(ctx)
    {
        if (!acc)
            return;
        if (c.timersUsed == c.timers.size())
        {
            lgpu_ctx::TimerSlot t;
            if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess)
                return;
            c.timers.push_back(t);
        }
        slot                = static_cast<int>(c.timersUsed++);
        c.timers[slot].acc  = acc;
        cudaEventRecord(c.timers[slot].a, c.stream);
    }
    ~StageTimer()
    {
        if (slot >= 0)
            cudaEventRecord(c.timers[slot].b, c.stream);
    }
};

// The reference computes alignment scores in int16 SIMD lanes (src/search_algo.hpp:1047,1087) and wraps silently beyond
// 32767; its answer is then meaningless, and we refuse to return a different one.  The scalar kernels (the only ones
// that can get there) raise the flag; checked once per call, after the stream has been synchronised.
static void checkScoreOverflow(lgpu_ctx & c)
{
    LGPU_CUDA(cudaMemcpyAsync(c.hOverflow.p, c.dOverflow.p, 4, cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaStreamSynchronize(c.stream));
    if (c.hOverflow.p[0])
    {
        LGPU_CUDA(cudaMemsetAsync(c.dOverflow.p, 0, 4, c.stream));
        throw UnsupportedError("an alignment score exceeds 32767: the reference computes in int16 lanes and wraps silently there "
                               "(src/search_algo.hpp:1047,1087); refusing to return a different answer");
    }
}

// call after the stream has been synchronised
static void resolveTimers(lgpu_ctx & c)
{
    for (size_t i = 0; i < c.timersUsed; ++i)
    {
        float ms = 0;
        if (c.timers[i].acc && cudaEventElapsedTime(&ms, c.timers[i].a, c.timers[i].b) == cudaSuccess)
            *c.timers[i].acc += ms;
        c.timers[i].acc = nullptr;
    }
    c.timersUsed = 0;
}

static void checkParams(lgpu_params const & p, lgpu_index_desc const & d)
{
    // same checks and messages as argConv0 (src/search.cpp:190-207)
    if (p.domain == LGPU_DOMAIN_PROTEIN)
    {
        if (d.trans_alph != LGPU_ALPH_AMINO_ACID)
            throw ArgError("Attempting to use nucleotide or bisulfite index for protein search.");
        if (d.orig_alph != LGPU_ALPH_AMINO_ACID && d.orig_alph != LGPU_ALPH_DNA5)
            throw ArgError("unknown original alphabet in index");
        if (p.query_alph != 0 && p.query_alph != LGPU_ALPH_AMINO_ACID && p.query_alph != LGPU_ALPH_DNA5)
            throw ArgError("query alphabet must be amino acid or dna5");
    }
    else if (p.domain == LGPU_DOMAIN_NUCLEOTIDE)
    {
        if (d.trans_alph != LGPU_ALPH_DNA5)
            throw ArgError("Attempting to use protein index for nucleotide search.");
        if (d.red_alph != LGPU_ALPH_DNA4)
            throw ArgError("Attempting to use bisulfite index for nucleotide search.");
        if (p.query_alph != 0 && p.query_alph != LGPU_ALPH_DNA5)
            throw ArgError("nucleotide searches take dna5 queries");
    }
    else if (p.domain == LGPU_DOMAIN_BISULFITE)
    {
        if (d.trans_alph != LGPU_ALPH_DNA5)
            throw ArgError("Attempting to use protein index for bisulfite search.");
        if (d.red_alph != LGPU_ALPH_DNA3BS)
            throw ArgError("Attempting to use nucleotid index for bisulfite search.");
        if (p.query_alph != 0 && p.query_alph != LGPU_ALPH_DNA5)
            throw ArgError("bisulfite searches take dna5 queries");
    }
    else
        throw ArgError("unknown domain");
    for (lgpu_search_opts const * o : {&p.opts0, &p.opts})
    {
        if (o->seed_length == 0 || o->seed_length > 2 * kMaxHalf2 || o->seed_offset == 0)
            throw ArgError("seed length must be in [1,32] and seed offset > 0");
        if (o->max_seed_dist > 2)
            throw UnsupportedError("seed distances > 2 are not implemented (reference default profiles use 0 or 1)");
        if (o->max_seed_dist >= 1 && !p.seed_half_exact && o->seed_length > kMaxHalf2)
            throw UnsupportedError("max_seed_dist > 0 without seed_half_exact supports seeds up to 16 residues");
    }
    if (p.max_matches == 0)
        throw ArgError("max_matches must be > 0");
}

// A batch as the pipeline sees it: residues in host or device memory, offsets on the host starting at 0.
struct BatchView
{
    uint8_t const *  residues    = nullptr;
    bool             resOnDevice = false;
    uint64_t const * offsets     = nullptr; // host, n + 1 entries, offsets[0] == 0
    uint64_t         n           = 0;
};

static void uploadQueries(lgpu_ctx & c, BatchView const & qb, lgpu_stats * st)
{
    lgpu_index const & ix = *c.index;
    c.nQueries            = qb.n;
    if (qb.n == 0)
        return;
    if (qb.n >= (1ull << 31) / c.di.qryNumFrames)
        throw ArgError("too many queries in one batch");
    c.qOffsHost.assign(qb.offsets, qb.offsets + qb.n + 1);
    if (c.qOffsHost[0] != 0)
        throw ArgError("query offsets must start at 0");
    c.totalResidues      = c.qOffsHost[qb.n];
    c.maxQueryLen        = 0;
    for (uint64_t i = 0; i < qb.n; ++i)
        c.maxQueryLen = std::max(c.maxQueryLen, static_cast<unsigned int>(c.qOffsHost[i + 1] - c.qOffsHost[i]));
    unsigned int const F = c.di.qryNumFrames;
    c.dQOrig.reserve(c.totalResidues);
    c.dQOffs.reserve(qb.n + 1);
    c.dQTrans.reserve(c.totalResidues * F);
    c.dQRed.reserve(c.totalResidues * F);
    {
        StageTimer t(c, st ? &st->ms_h2d : nullptr);
        LGPU_CUDA(cudaMemcpyAsync(c.dQOrig.p, qb.residues, c.totalResidues,
                                  qb.resOnDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c.stream));
        LGPU_CUDA(cudaMemcpyAsync(c.dQOffs.p, c.qOffsHost.data(), (qb.n + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    }
    DevQueries & Q = c.Q;
    Q.orig         = c.dQOrig.p;
    Q.offs         = c.dQOffs.p;
    Q.trans        = c.dQTrans.p;
    Q.red          = c.dQRed.p;
    Q.n            = static_cast<unsigned int>(qb.n);
    Q.F            = F;
    Q.frameMode    = c.di.qFrameMode;
    std::memset(Q.redTab, 0, sizeof(Q.redTab));
    std::memset(Q.compTab, 0, sizeof(Q.compTab));
    switch (ix.meta.red_alph)
    {
        case LGPU_ALPH_LI10: std::memcpy(Q.redTab[0], kAa27ToLi10, 27); break;
        case LGPU_ALPH_MURPHY10: std::memcpy(Q.redTab[0], kAa27ToMurphy10, 27); break;
        case LGPU_ALPH_AMINO_ACID:
            for (int i = 0; i < 27; ++i)
                Q.redTab[0][i] = static_cast<unsigned char>(i);
            break;
        case LGPU_ALPH_DNA4:
        {
            // dna5 (A,C,G,N,T) -> dna4 (A,C,G,T).  The reference replaces every READ of an N by the next output of a
            // per-view random generator: N is stored as a marker and resolved by the seeding
            // kernels exactly like that (n_random.hpp, seedSym / elongSym).
            unsigned char const t[5] = {0, 1, 2, static_cast<unsigned char>(kNMarker), 3};
            std::memcpy(Q.redTab[0], t, 5);
            break;
        }
        case LGPU_ALPH_DNA3BS:
        {
            // dna5 -> dna4 (N -> marker | direction, as above) -> bisulfite semialphabet: forward (C->T) ranks
            // {A0,C1,G2,T1} for even frames, reverse (G->A) ranks {A3,C4,G3,T5} for odd frames
            // (src/view_reduce_to_bisulfite.hpp:51-52)
            unsigned char const fwd[5] = {0, 1, 2, static_cast<unsigned char>(kNMarker), 1},
                                rev[5] = {3, 4, 3, static_cast<unsigned char>(kNMarker | 1u), 5};
            std::memcpy(Q.redTab[0], fwd, 5);
            std::memcpy(Q.redTab[1], rev, 5);
            break;
        }
        default: throw UnsupportedError("unsupported reduced alphabet");
    }
    if (ix.meta.red_alph != LGPU_ALPH_DNA3BS)
        std::memcpy(Q.redTab[1], Q.redTab[0], 32);
    std::memcpy(Q.compTab, kDna5Complement, 5);
    prepQueriesKernel<<<std::min<unsigned int>(Q.n, 65535u * 8u), 128, 0, c.stream>>>(Q);
    LGPU_CUDA(cudaGetLastError());
    if (st)
        st->kernel_launches += 1;
}

// Depth of the prefix table for a seed whose exactly matched part has h1 symbols: as deep as 2^22 entries (32 MB,
// L2-resident next to the packed occurrence blocks of a short-read index) allow.
static unsigned int prefixDepth(lgpu_index const & ix, unsigned int h1)
{
    uint64_t const A = ix.dev.sigma - 1;
    if (A < 2 || ix.dev.nRows >= (1ull << 32))
        return 0;
    unsigned int k = 0;
    uint64_t     n = 1;
    while (k < h1 && n * A <= (1ull << 22))
    {
        n *= A;
        ++k;
    }
    return k >= 2 ? k : 0;
}

static uint2 const * prefixTable(lgpu_index const & ix, unsigned int k, cudaStream_t stream)
{
    if (k == 0)
        return nullptr;
    std::lock_guard<std::mutex> lock(ix.prefixMutex);
    auto it = ix.prefixTabs.find(k);
    if (it != ix.prefixTabs.end())
        return it->second;
    uint64_t n = 1;
    for (unsigned int i = 0; i < k; ++i)
        n *= ix.dev.sigma - 1;
    uint2 * tab = nullptr;
    LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&tab), n * sizeof(uint2)));
    prefixTableKernel<<<gridFor(n, 128), 128, 0, stream>>>(ix.dev, k, static_cast<unsigned int>(n), tab);
    LGPU_CUDA(cudaGetLastError());
    LGPU_CUDA(cudaStreamSynchronize(stream));
    ix.prefixTabs.emplace(k, tab);
    return tab;
}

// search(): returns number of matches now in c.dMatches
static uint64_t runSeeding(lgpu_ctx & c, lgpu_search_opts const & so, unsigned int const * dActive, unsigned int nActive,
                           unsigned int maxActiveLen, lgpu_stats * st)
{
    if (nActive == 0)
        return 0;
    StageTimer t(c, st ? &st->ms_seed : nullptr);
    c.dCounters.reserve(8);
    size_t cap = std::max<size_t>(c.dMatches.cap, std::max<size_t>(1u << 20, static_cast<size_t>(nActive) * 64));
    for (int attempt = 0; attempt < 3; ++attempt)
    {
        c.dMatches.reserve(cap);
        LGPU_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 8 * sizeof(unsigned long long), c.stream));
        SeedParams P;
        P.ix               = c.index->dev;
        P.Q                = c.Q;
        P.active           = dActive;
        P.nActive          = nActive;
        P.seedLength       = so.seed_length;
        P.seedOffset       = so.seed_offset;
        P.maxSeedDist      = so.max_seed_dist;
        P.halfExact        = c.params.seed_half_exact;
        P.fullHamming      = (so.max_seed_dist >= 1 && !c.params.seed_half_exact) ? 1u : 0u;
        P.adaptive         = c.params.adaptive_seeding;
        P.maxMatches       = c.params.max_matches;
        P.preScoring       = c.params.pre_scoring;
        P.preScoringThresh = c.params.pre_scoring_thresh;
        P.unknownRank      = c.di.unknownRank;
        P.matrix           = c.dMatrix.p;
        P.out              = c.dMatches.p;
        P.cap              = c.dMatches.cap;
        P.counters         = c.dCounters.p;
        {
            bool const         halfM = so.max_seed_dist != 0 && c.params.seed_half_exact;
            unsigned int const h1    = P.fullHamming ? 0u : (halfM ? so.seed_length / 2 : so.seed_length);
            P.prefixK                = c.seedPrefix ? prefixDepth(*c.index, h1) : 0u;
            P.prefixTab              = prefixTable(*c.index, P.prefixK, c.stream);
            P.textElong              = c.seedTextElong ? 1u : 0u;
            for (int par = 0; par < 2; ++par)
                for (int r = 0; r < 32; ++r)
                    P.sbjRed[par][r] = c.Q.redTab[par][r] >= kNMarker ? 0xffu : c.Q.redTab[par][r];
        }
        // Few queries (typically the phase-2 leftovers): latency matters -> one block per query, the seeds
        // of a query searched by 16 warps in parallel.  Many queries: throughput matters -> one warp per
        // query for half-exact seeds (wide search tree per seed), one thread per query for exact seeds
        // (most independent chains in flight).
        bool const         halfMode  = so.max_seed_dist != 0 && c.params.seed_half_exact;
        unsigned int const n2        = P.fullHamming ? so.seed_length : (halfMode ? so.seed_length - so.seed_length / 2 : 0);
        unsigned int const A         = c.index->dev.sigma - 1;
        unsigned int const maxLeaves = (A - 1) * n2 + 1 + (so.max_seed_dist >= 2 ? n2 * (n2 - 1) / 2 * (A - 1) * (A - 1) : 0u);
        unsigned int const maxSeeds  = maxActiveLen >= so.seed_length
                                         ? c.di.qryNumFrames * ((maxActiveLen - so.seed_length) / so.seed_offset + 1)
                                         : 1;
        size_t const scratchCursors = static_cast<size_t>(nActive) * maxSeeds * maxLeaves;
        bool const   blockPerQuery  = c.seedMode != 1 && c.seedMode != 2 && c.seedMode != 4 && nActive <= 4096 &&
                                   maxSeeds <= static_cast<unsigned int>(kSeedBlockMaxSeeds) && maxActiveLen < (1u << 24) &&
                                   scratchCursors * sizeof(Cursor) <= (1ull << 30);
        if (blockPerQuery || c.seedMode == 3)
        {
            if (!blockPerQuery)
                throw ArgError("LAMBDA_B200_SEED=block: batch too large for the block-per-query kernel");
            c.dSeedCursors.reserve(scratchCursors * 2);
            c.dSeedCounts.reserve(static_cast<size_t>(nActive) * maxSeeds);
            SeedScratch S;
            S.cursors   = reinterpret_cast<Cursor *>(c.dSeedCursors.p);
            S.counts    = c.dSeedCounts.p;
            S.maxSeeds  = maxSeeds;
            S.maxLeaves = maxLeaves;
            seedBlockKernel<<<nActive, 32 * kSeedBlockWarps, 0, c.stream>>>(P, S);
        }
        else if (c.seedMode == 2)
            seedWarpKernel<<<gridFor(static_cast<unsigned long long>(nActive) * 32, 128), 128, 0, c.stream>>>(P);
        else if (c.seedMode == 1 && !P.fullHamming && so.max_seed_dist < 2) // the thread-per-query kernel: no level order, one mismatch
            seedKernel<<<gridFor(nActive, 128), 128, 0, c.stream>>>(P);
        else
            seedSpecKernel<<<gridFor(static_cast<unsigned long long>(nActive) * 32, 32 * kSpecWarps), 32 * kSpecWarps, 0,
                             c.stream>>>(P);
        LGPU_CUDA(cudaGetLastError());
        if (st)
            st->kernel_launches += 1;
        unsigned long long cnt[3];
        LGPU_CUDA(cudaMemcpyAsync(cnt, c.dCounters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c.stream));
        syncStream(c);
        if (cnt[0] <= c.dMatches.cap)
        {
            if (st)
            {
                st->hits_after_seeding += cnt[1];
                st->hits_failed_pre_extend += cnt[2];
            }
            return cnt[0];
        }
        cap = cnt[0]; // the buffer was too small: the count is exact, so one retry suffices
    }
    throw CudaError("seeding output buffer overflow");
}

// _widenAndPreprocessMatches on dIn[0..n); result in c.dMerged, returns the merged count
static uint64_t runMerge(lgpu_ctx & c, lgpu_match const * dIn, uint64_t n, lgpu_stats * st)
{
    if (n == 0)
        return 0;
    if (n >= (1ull << 31)) // the CUB primitives below take int item counts
        throw ArgError("more than 2^31 seed matches in one batch; use smaller query batches");
    StageTimer t(c, st ? &st->ms_sort_merge : nullptr);
    c.dKey1.reserve(n);
    c.dKey2.reserve(n);
    c.dKey1b.reserve(n);
    c.dKey2b.reserve(n);
    c.dPerm.reserve(n);
    c.dPermB.reserve(n);
    c.dHead.reserve(n);
    c.dScan.reserve(n);
    unsigned int const g = gridFor(n, 256);
    widenKernel<<<g, 256, 0, c.stream>>>(dIn, n, c.Q, c.index->dev, c.params.window_band, c.dKey1.p, c.dKey2.p);
    iotaKernel<<<g, 256, 0, c.stream>>>(c.dPerm.p, n);
    // LSD: stable sort by the minor key (window start/end), then by the major key (qry, subj)
    size_t tmp1 = 0, tmp2 = 0, tmp3 = 0;
    int const nI = static_cast<int>(n);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp1, c.dKey2.p, c.dKey2b.p, c.dPerm.p, c.dPermB.p, nI, 0, 64, c.stream);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp2, c.dKey1b.p, c.dKey1.p, c.dKey2b.p, c.dKey2.p, nI, 0, 64, c.stream);
    cub::DeviceScan::InclusiveSum(nullptr, tmp3, c.dHead.p, c.dScan.p, nI, c.stream);
    c.dCubTemp.reserve(std::max(tmp1, std::max(tmp2, tmp3)));
    size_t tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceRadixSort::SortPairs(c.dCubTemp.p, tb, c.dKey2.p, c.dKey2b.p, c.dPerm.p, c.dPermB.p, nI, 0, 64,
                                              c.stream));
    gatherKernel<<<g, 256, 0, c.stream>>>(c.dKey1.p, c.dPermB.p, n, c.dKey1b.p); // key1 in minor-sorted order
    tb = c.dCubTemp.cap;
    // sort (key1b -> key1) carrying key2b -> key2
    LGPU_CUDA(cub::DeviceRadixSort::SortPairs(c.dCubTemp.p, tb, c.dKey1b.p, c.dKey1.p, c.dKey2b.p, c.dKey2.p, nI, 0, 64,
                                              c.stream));
    chainHeadKernel<<<g, 256, 0, c.stream>>>(c.dKey1.p, c.dKey2.p, n, c.dHead.p);
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::InclusiveSum(c.dCubTemp.p, tb, c.dHead.p, c.dScan.p, nI, c.stream));
    unsigned int nChains = 0;
    LGPU_CUDA(cudaMemcpyAsync(&nChains, c.dScan.p + (n - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    c.dMerged.reserve(nChains);
    chainEmitKernel<<<g, 256, 0, c.stream>>>(c.dKey1.p, c.dKey2.p, c.dHead.p, c.dScan.p, n, c.Q, c.dMerged.p);
    LGPU_CUDA(cudaGetLastError());
    if (st)
    {
        st->kernel_launches += 5 + 3; // ours + the three CUB primitives
        st->hits_duplicate += n - nChains;
    }
    return nChains;
}

// columns per lane of the wavefront kernel: the smallest even K with 32 * K >= longest query, capped at 16
static int chooseK(unsigned int maxQ)
{
    unsigned int const k = (maxQ + 31) / 32;
    if (k <= 2) return 2;
    if (k >= 16) return 16;
    return static_cast<int>((k + 1) / 2 * 2);
}

template <bool TRACE>
static void launchWavefront(int K, ExtParams const & P, unsigned int grid, cudaStream_t s)
{
    switch (K)
    {
        case 2: swWavefrontKernel<2, TRACE><<<grid, 128, 0, s>>>(P); break;
        case 4: swWavefrontKernel<4, TRACE><<<grid, 128, 0, s>>>(P); break;
        case 6: swWavefrontKernel<6, TRACE><<<grid, 128, 0, s>>>(P); break;
        case 8: swWavefrontKernel<8, TRACE><<<grid, 128, 0, s>>>(P); break;
        case 10: swWavefrontKernel<10, TRACE><<<grid, 128, 0, s>>>(P); break;
        case 12: swWavefrontKernel<12, TRACE><<<grid, 128, 0, s>>>(P); break;
        case 14: swWavefrontKernel<14, TRACE><<<grid, 128, 0, s>>>(P); break;
        default: swWavefrontKernel<16, TRACE><<<grid, 128, 0, s>>>(P); break;
    }
}

struct TaskDims
{
    unsigned int maxQ = 0, maxT = 0;
    uint64_t     cells = 0;
};

// dimensions of a host-side task list (trace-buffer layout, work statistics)
static TaskDims taskDims(lgpu_match const * tasks, size_t n)
{
    TaskDims d;
    for (size_t i = 0; i < n; ++i)
    {
        unsigned int const nq = tasks[i].qry_end - tasks[i].qry_start, nt = tasks[i].subj_end - tasks[i].subj_start;
        d.maxQ                = std::max(d.maxQ, nq);
        d.maxT                = std::max(d.maxT, nt);
        d.cells += static_cast<uint64_t>(nq) * nt;
    }
    return d;
}

static ExtParams baseExtParams(lgpu_ctx & c, lgpu_match const * dTasks, unsigned int n, unsigned int grid, TaskDims const & dims,
                               int K)
{
    ExtParams P{};
    P.ix          = c.index->dev;
    P.Q           = c.Q;
    P.tasks       = dTasks;
    P.nTasks      = n;
    P.matrix      = c.dMatrix.p;
    P.go          = c.scoring.gapOpenSeqan;
    P.ge          = c.scoring.gapExtend;
    c.dWork.reserve(kMaxDpxClasses + 3);
    P.workCounter = c.dWork.p;
    P.overflowFlag = c.dOverflow.p;
    P.order       = nullptr;
    P.maxRows     = std::max(dims.maxT, 1u);
    if (dims.maxQ > static_cast<unsigned int>(32 * K))
        c.dBoundary.reserve(static_cast<size_t>(grid) * 4 * P.maxRows);
    else
        c.dBoundary.reserve(1);
    P.boundary = c.dBoundary.p;
    return P;
}

template <int T, int K, bool PRIV, bool TRACE>
static void launchDpx(lgpu_ctx & c, DpxParams P, unsigned int maxNt, cudaStream_t stream)
{
    P.winCap          = (maxNt + 4 * T + 127) / 128 * 128;
    size_t const smem = dpxSmemBytes(T, K, PRIV, P.nCodes, P.winCap);
    if (smem > 227 * 1024)
        throw CudaError("DPX kernel: window too long for shared memory");
    LGPU_CUDA(cudaFuncSetAttribute(swDpxKernel<T, K, PRIV, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    unsigned int const perSM = TRACE ? 32u : c.dpxBlocksPerSM;
    unsigned int const grid  = std::min<unsigned int>(P.nJobs, static_cast<unsigned int>(c.numSMs) * perSM);
    swDpxKernel<T, K, PRIV, TRACE><<<grid, 32, smem, stream>>>(P);
    LGPU_CUDA(cudaGetLastError());
}

struct MaxOp
{
    __host__ __device__ unsigned int operator()(unsigned int a, unsigned int b) const { return a > b ? a : b; }
};

using DpxLaunchFn = void (*)(lgpu_ctx &, DpxParams, unsigned int, cudaStream_t);
#define LGPU_DPX_LAUNCH_SHARED(T, K) &launchDpx<T, K, false, false>,
#define LGPU_DPX_LAUNCH_PRIV(T, K) &launchDpx<T, K, true, false>,
#define LGPU_DPX_LAUNCH_TRACE(T, K) &launchDpx<T, K, true, true>,
static DpxLaunchFn const kDpxLaunchShared[kNumDpxClasses]      = {LGPU_DPX_CLASSES(LGPU_DPX_LAUNCH_SHARED)};
static DpxLaunchFn const kDpxLaunchPriv[kNumDpxPrivClasses]    = {LGPU_DPX_PRIV_CLASSES(LGPU_DPX_LAUNCH_PRIV)};
static DpxLaunchFn const kDpxLaunchPrivTr[kNumDpxPrivClasses]  = {LGPU_DPX_PRIV_CLASSES(LGPU_DPX_LAUNCH_TRACE)};
static DpxLaunchFn const kDpxLaunchTrace32[kNumDpxTr32Classes] = {LGPU_DPX_TRACE32_CLASSES(LGPU_DPX_LAUNCH_TRACE)};
#undef LGPU_DPX_LAUNCH_SHARED
#undef LGPU_DPX_LAUNCH_PRIV
#undef LGPU_DPX_LAUNCH_TRACE

constexpr int kClsSlots = kMaxDpxClasses + 1;

// The (T, K) classes of a pass are independent launches.  With many classes of few alignments each (real length
// distributions: 31 classes for 164 k alignments) one launch after the other leaves most SMs idle at the tail of
// every class, so the classes go round-robin onto a few side streams: callers launch after the host has waited for
// the context's stream (all inputs are ready) and call joinClassStreams() before anything that consumes the results.
static cudaStream_t classStream(lgpu_ctx & c, unsigned int k, unsigned int nClasses)
{
    if (nClasses < 3 || !c.classStreams)
        return c.stream;
    if (!c.aux[0])
        for (int i = 0; i < lgpu_ctx::kAux; ++i)
        {
            LGPU_CUDA(cudaStreamCreateWithFlags(&c.aux[i], cudaStreamNonBlocking));
            LGPU_CUDA(cudaEventCreateWithFlags(&c.auxDone[i], cudaEventDisableTiming));
        }
    unsigned int const slot = k % (lgpu_ctx::kAux + 1);
    if (slot == 0)
        return c.stream;
    c.auxUsed |= 1u << (slot - 1);
    return c.aux[slot - 1];
}

static void joinClassStreams(lgpu_ctx & c)
{
    for (int i = 0; i < lgpu_ctx::kAux; ++i)
        if (c.auxUsed & (1u << i))
        {
            LGPU_CUDA(cudaEventRecord(c.auxDone[i], c.aux[i]));
            LGPU_CUDA(cudaStreamWaitEvent(c.stream, c.auxDone[i], 0));
        }
    c.auxUsed = 0;
} // per-class arrays: the packed classes of a table + the scalar class

static DpxParams baseDpxParams(lgpu_ctx & c, lgpu_match const * dTasks, unsigned int n)
{
    DpxParams P{};
    P.ix      = c.index->dev;
    P.Q       = c.Q;
    P.tasks   = dTasks;
    P.nSorted = n;
    P.matrix  = c.dMatrix.p;
    P.go      = c.scoring.gapOpenSeqan;
    P.ge      = c.scoring.gapExtend;
    P.nCodes  = static_cast<unsigned int>(c.scoring.alphSize) + 1;
    return P;
}

// DP pass 1: scores into dScores[0..n).  Alignments are sorted by length class of the query, then -- shared
// profiles -- by (query, window length): up to 32/T consecutive alignments of one query form a job of the packed
// kernel; or -- private profiles (small alphabets) -- by window length alone: any 32/T consecutive alignments form
// a job.  Whatever does not fit the packed kernel (queries > 2048, windows > 8192, exotic scoring) runs on the scalar
// wavefront kernel.
static void runScorePass(lgpu_ctx & c, lgpu_match const * dTasks, unsigned int n, int * dScores, lgpu_stats * st)
{
    if (n == 0)
        return;
    if (n >= (1u << 31))
        throw ArgError("more than 2^31 alignments in one batch; use smaller query batches");
    StageTimer t(c, st ? &st->ms_extend_score : nullptr);
    bool const priv = c.privProfiles;
    int const  tab  = priv ? kDpxTabPriv : kDpxTabShared;
    int const  NP   = dpxNumClasses(tab); // packed classes; class NP = scalar
    constexpr int NC = kClsSlots;
    c.dClassKeys.reserve(n);
    c.dClassKeysB.reserve(n);
    c.dOrder.reserve(n);
    c.dOrderB.reserve(n);
    c.dClassInfo.reserve(3 * NC + 4);
    c.dCounters.reserve(8);
    c.dWork.reserve(NC + 1);
    LGPU_CUDA(cudaMemsetAsync(c.dClassInfo.p, 0, (3 * NC + 4) * 4, c.stream));
    LGPU_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 8 * sizeof(unsigned long long), c.stream));
    LGPU_CUDA(cudaMemsetAsync(c.dWork.p, 0, (NC + 1) * 4, c.stream));
    unsigned int const g  = gridFor(n, 256);
    int const          nI = static_cast<int>(n);
    classifyKernel<<<g, 256, 0, c.stream>>>(dTasks, n, c.index->dev.bsMode, tab, c.maxColsInt16, c.dClassKeys.p, c.dOrder.p,
                                            c.dClassInfo.p, c.dClassInfo.p + NC, c.dClassInfo.p + 3 * NC, c.dCounters.p, nullptr);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, c.dClassKeys.p, c.dClassKeysB.p, c.dOrder.p, c.dOrderB.p, nI, 0, 64, c.stream);
    if (!priv)
    {
        c.dSegStart.reserve(n);
        c.dSegStartB.reserve(n);
        c.dJobHead.reserve(n);
        c.dJobPos.reserve(n);
        c.dJobs.reserve(n);
        cub::DeviceScan::InclusiveScan(nullptr, t2, c.dSegStart.p, c.dSegStartB.p, MaxOp(), nI, c.stream);
        cub::DeviceScan::InclusiveSum(nullptr, t3, c.dJobHead.p, c.dJobPos.p, nI, c.stream);
    }
    c.dCubTemp.reserve(std::max(t1, std::max(t2, t3)));
    size_t tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceRadixSort::SortPairs(c.dCubTemp.p, tb, c.dClassKeys.p, c.dClassKeysB.p, c.dOrder.p, c.dOrderB.p, nI, 0, 64,
                                              c.stream));
    unsigned int launches = 1 + 1;
    if (!priv)
    {
        segFlagKernel<<<g, 256, 0, c.stream>>>(c.dClassKeysB.p, c.dOrderB.p, dTasks, n, c.dSegStart.p);
        tb = c.dCubTemp.cap;
        LGPU_CUDA(cub::DeviceScan::InclusiveScan(c.dCubTemp.p, tb, c.dSegStart.p, c.dSegStartB.p, MaxOp(), nI, c.stream));
        jobHeadKernel<<<g, 256, 0, c.stream>>>(c.dClassKeysB.p, c.dSegStartB.p, n, c.dJobHead.p, c.dClassInfo.p + 2 * NC);
        tb = c.dCubTemp.cap;
        LGPU_CUDA(cub::DeviceScan::InclusiveSum(c.dCubTemp.p, tb, c.dJobHead.p, c.dJobPos.p, nI, c.stream));
        jobEmitKernel<<<g, 256, 0, c.stream>>>(c.dJobHead.p, c.dJobPos.p, n, c.dJobs.p);
        launches += 5;
    }
    LGPU_CUDA(cudaGetLastError());
    unsigned int       info[3 * NC + 1];
    unsigned long long cells = 0;
    LGPU_CUDA(cudaMemcpyAsync(info, c.dClassInfo.p, sizeof(info), cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&cells, c.dCounters.p, 8, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    unsigned int taskOff = 0, jobOff = 0, nNonEmpty = 0, kLaunch = 0;
    for (int cls = 0; cls <= NP; ++cls)
        nNonEmpty += info[cls] ? 1u : 0u;
    for (int cls = 0; cls <= NP; ++cls)
    {
        unsigned int const cnt = info[cls], maxNt = info[NC + cls];
        if (cnt == 0)
            continue;
        bool const         scalar = (cls == NP) || !c.dpxScoreOk;
        cudaStream_t const cs     = scalar ? c.stream : classStream(c, kLaunch++, nNonEmpty);
        unsigned int const nJobs  = priv ? (cnt + dpxGroupsOf(tab, cls) - 1) / dpxGroupsOf(tab, cls) : info[2 * NC + cls];
        if (!scalar)
        {
            DpxParams P   = baseDpxParams(c, dTasks, n);
            P.order       = c.dOrderB.p;
            P.keys        = c.dClassKeysB.p;
            P.jobs        = priv ? nullptr : c.dJobs.p + jobOff;
            P.nJobs       = nJobs;
            P.slotBase    = taskOff;
            P.nSlots      = cnt;
            P.workCounter = c.dWork.p + cls;
            P.scores      = dScores;
            (priv ? kDpxLaunchPriv[cls] : kDpxLaunchShared[cls])(c, P, maxNt, cs);
        }
        else
        {
            TaskDims d2;
            d2.maxQ = info[3 * NC];
            d2.maxT = maxNt;
            int const          K    = chooseK(d2.maxQ);
            unsigned int const grid = std::min<unsigned int>((cnt + 3) / 4, static_cast<unsigned int>(c.numSMs) * 16);
            ExtParams          P    = baseExtParams(c, dTasks, cnt, grid, d2, K);
            P.order                 = c.dOrderB.p + taskOff;
            P.workCounter           = c.dWork.p + cls;
            P.scores                = dScores;
            launchWavefront<false>(K, P, grid, c.stream);
            LGPU_CUDA(cudaGetLastError());
        }
        ++launches;
        taskOff += cnt;
        jobOff += priv ? 0u : nJobs;
    }
    joinClassStreams(c);
    if (st)
    {
        st->kernel_launches += launches;
        st->n_extensions_score += n;
        st->cells_score += cells;
    }
}

// Second traceback pass for lgpu_params.want_cigar: the first pass left n_gap_open in c.dHits, which bounds the
// number of runs of every alignment; `launch(emitParams...)` re-walks the same trace and writes the runs.
// `order` = device list of the `cnt` hit slots handled by this launch (nullptr: slots 0 .. cnt-1).
template <typename TLaunch>
static void emitCigars(lgpu_ctx & c, unsigned int const * order, unsigned int cnt, TLaunch && launch, lgpu_stats * st)
{
    if (cnt == 0)
        return;
    c.dCigarCap.reserve(cnt + 1);
    c.dCigarOff.reserve(cnt + 1);
    cigarCapKernel<<<gridFor(cnt, 256), 256, 0, c.stream>>>(c.dHits.p, order, cnt, c.dCigarCap.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, c.dCigarCap.p, c.dCigarOff.p, static_cast<int>(cnt), c.stream);
    c.dCubTemp.reserve(tb);
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::ExclusiveSum(c.dCubTemp.p, tb, c.dCigarCap.p, c.dCigarOff.p, static_cast<int>(cnt), c.stream));
    unsigned int lastOff = 0, lastCap = 0;
    LGPU_CUDA(cudaMemcpyAsync(&lastOff, c.dCigarOff.p + (cnt - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&lastCap, c.dCigarCap.p + (cnt - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    size_t const total = static_cast<size_t>(lastOff) + lastCap;
    if (c.cigar.size() + total > 0xffffffffull)
        throw ArgError("too many alignment operations in one batch; use smaller batches with want_cigar");
    c.dCigar.reserve(total);
    c.hCigarStage.reserve(total);
    launch(c.dCigar.p, c.dCigarOff.p, static_cast<unsigned int>(c.cigar.size()));
    LGPU_CUDA(cudaGetLastError());
    LGPU_CUDA(cudaMemcpyAsync(c.hCigarStage.p, c.dCigar.p, total * 4, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    c.cigar.insert(c.cigar.end(), c.hCigarStage.p, c.hCigarStage.p + total);
    if (st)
        st->kernel_launches += 3;
}

// DP pass 2 + traceback on the scalar wavefront kernel (1 trace byte per cell) for `n` tasks (host copy `tasks`,
// device copy dTasks); the record of task t lands in c.dHits[outIdx[t]] (outIdx = nullptr: c.dHits[t])
static void runTraceScalar(lgpu_ctx & c, lgpu_match const * tasks, size_t n, lgpu_match const * dTasks, unsigned int const * dOutIdx,
                           lgpu_stats * st)
{
    TaskDims const dims = taskDims(tasks, n);
    int const      K    = chooseK(dims.maxQ);
    unsigned int const cols = 32 * K, colBytes = 32 * ((K + 3) / 4 * 4);
    // chunk so that the trace matrices of one launch stay below 16 GiB of the 180 GB HBM
    constexpr uint64_t kMaxTraceBytes = 16ull << 30;
    std::vector<unsigned long long> offs(n);
    c.dScores2.reserve(n);
    c.dBestPos.reserve(2 * n);
    c.dTraceOff.reserve(n);
    size_t begin = 0;
    while (begin < n)
    {
        uint64_t bytes = 0;
        size_t   end   = begin;
        while (end < n)
        {
            unsigned int const nq = tasks[end].qry_end - tasks[end].qry_start, nt = tasks[end].subj_end - tasks[end].subj_start;
            uint64_t const     sz = static_cast<uint64_t>((nq + cols - 1) / cols * colBytes) * nt;
            if (end > begin && bytes + sz > kMaxTraceBytes)
                break;
            offs[end] = bytes;
            bytes += (sz + 15) / 16 * 16;
            ++end;
        }
        unsigned int const cnt = static_cast<unsigned int>(end - begin);
        c.dTrace.reserve(bytes);
        LGPU_CUDA(cudaMemcpyAsync(c.dTraceOff.p, offs.data() + begin, cnt * 8ull, cudaMemcpyHostToDevice, c.stream));
        unsigned int const grid = std::min<unsigned int>((cnt + 3) / 4, static_cast<unsigned int>(c.numSMs) * 16);
        ExtParams          P    = baseExtParams(c, dTasks + begin, cnt, grid, dims, K);
        P.scores                = c.dScores2.p;
        P.bestPos               = c.dBestPos.p;
        P.trace                 = c.dTrace.p;
        P.traceOff              = c.dTraceOff.p;
        LGPU_CUDA(cudaMemsetAsync(c.dWork.p, 0, 4, c.stream));
        launchWavefront<true>(K, P, grid, c.stream);
        LGPU_CUDA(cudaGetLastError());
        TracebackParams TP{};
        TP.ix           = c.index->dev;
        TP.Q            = c.Q;
        TP.tasks        = dTasks + begin;
        TP.nTasks       = cnt;
        TP.matrix       = c.dMatrix.p;
        TP.scores       = c.dScores2.p;
        TP.bestPos      = c.dBestPos.p;
        TP.trace        = c.dTrace.p;
        TP.traceOff     = c.dTraceOff.p;
        TP.K            = static_cast<unsigned int>(K);
        TP.out          = c.dHits.p + (dOutIdx ? 0 : begin);
        TP.outIndex     = dOutIdx ? dOutIdx + begin : nullptr;
        tracebackKernel<<<gridFor(cnt, 128), 128, 0, c.stream>>>(TP);
        LGPU_CUDA(cudaGetLastError());
        if (c.params.want_cigar)
        {
            // the emit pass addresses the records through a slot list
            c.dScalarSlots.reserve(cnt);
            if (dOutIdx)
                LGPU_CUDA(cudaMemcpyAsync(c.dScalarSlots.p, dOutIdx + begin, cnt * 4ull, cudaMemcpyDeviceToDevice, c.stream));
            else
                iotaFromKernel<<<gridFor(cnt, 256), 256, 0, c.stream>>>(c.dScalarSlots.p, cnt, static_cast<unsigned int>(begin));
            TracebackParams TE = TP;
            TE.out             = c.dHits.p;
            TE.outIndex        = c.dScalarSlots.p;
            emitCigars(c, c.dScalarSlots.p, cnt, [&](unsigned int * ops, unsigned int const * off, unsigned int base) {
                TE.emit      = 1;
                TE.cigarOps  = ops;
                TE.cigarOff  = off;
                TE.cigarBase = base;
                tracebackKernel<<<gridFor(cnt, 128), 128, 0, c.stream>>>(TE);
            }, st);
        }
        syncStream(c); // offs / the trace buffer are reused by the next chunk
        if (st)
            st->kernel_launches += 2;
        begin = end;
    }
}

// DP pass 2 + traceback for the `n` alignments dTasks[0..n) (device); the record of task t lands in c.dHits[t].
// Everything is laid out on the device: the alignments are classified and sorted by (class, window length), a scan of
// the plane sizes gives every alignment its residue plane, one fill launch per class and ONE traceback launch walk
// them.  The host reads back the class counts once (launch configuration, plane buffer size).  Alignments outside
// the packed kernels' range (queries > 2048 columns, windows > 8192, exotic scoring) take the scalar path.
static void runTracePass(lgpu_ctx & c, lgpu_match const * dTasks, unsigned int n, lgpu_stats * st)
{
    if (n == 0)
        return;
    if (n >= (1u << 31))
        throw ArgError("more than 2^31 alignments in one batch; use smaller query batches");
    StageTimer t(c, st ? &st->ms_extend_trace : nullptr);
    c.dHits.reserve(n);
    c.dWork.reserve(kClsSlots + 2);
    if (!c.resTraceOk || c.forceScalarTrace)
    {
        // no packed path for this scoring scheme: everything on the scalar kernel
        std::vector<lgpu_match> host(n);
        LGPU_CUDA(cudaMemcpyAsync(host.data(), dTasks, n * sizeof(lgpu_match), cudaMemcpyDeviceToHost, c.stream));
        syncStream(c);
        runTraceScalar(c, host.data(), n, dTasks, nullptr, st);
        if (st)
        {
            st->n_extensions_trace += n;
            st->cells_trace += taskDims(host.data(), n).cells;
        }
        return;
    }
    int const     tab = c.traceTab;
    int const     NP  = dpxNumClasses(tab);
    constexpr int NC  = kClsSlots;
    c.dClassKeys.reserve(n);
    c.dClassKeysB.reserve(n);
    c.dOrder.reserve(n);
    c.dOrderB.reserve(n);
    c.dPlaneWords.reserve(n);
    c.dPlaneWordsB.reserve(n);
    c.dTraceOff.reserve(n);
    c.dScores2.reserve(n);
    c.dBestPos.reserve(n);
    c.dClassInfo.reserve(3 * NC + 4);
    c.dCounters.reserve(8);
    LGPU_CUDA(cudaMemsetAsync(c.dClassInfo.p, 0, (3 * NC + 4) * 4, c.stream));
    LGPU_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 8 * sizeof(unsigned long long), c.stream));
    LGPU_CUDA(cudaMemsetAsync(c.dWork.p, 0, (NC + 1) * 4, c.stream));
    unsigned int const g  = gridFor(n, 256);
    int const          nI = static_cast<int>(n);
    classifyKernel<<<g, 256, 0, c.stream>>>(dTasks, n, c.index->dev.bsMode, tab, c.maxColsInt16, c.dClassKeys.p, c.dOrder.p,
                                            c.dClassInfo.p, c.dClassInfo.p + NC, c.dClassInfo.p + 3 * NC, c.dCounters.p,
                                            c.dPlaneWords.p);
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, c.dClassKeys.p, c.dClassKeysB.p, c.dOrder.p, c.dOrderB.p, nI, 0, 64, c.stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, c.dPlaneWordsB.p, c.dTraceOff.p, nI, c.stream);
    c.dCubTemp.reserve(std::max(t1, t2));
    size_t tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceRadixSort::SortPairs(c.dCubTemp.p, tb, c.dClassKeys.p, c.dClassKeysB.p, c.dOrder.p, c.dOrderB.p, nI, 0, 64,
                                              c.stream));
    gatherU64Kernel<<<g, 256, 0, c.stream>>>(c.dPlaneWords.p, c.dOrderB.p, n, c.dPlaneWordsB.p);
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::ExclusiveSum(c.dCubTemp.p, tb, c.dPlaneWordsB.p, c.dTraceOff.p, nI, c.stream));
    LGPU_CUDA(cudaGetLastError());
    unsigned int       info[3 * NC + 1];
    unsigned long long cells = 0, lastOff = 0, lastWords = 0;
    LGPU_CUDA(cudaMemcpyAsync(info, c.dClassInfo.p, sizeof(info), cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&cells, c.dCounters.p, 8, cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&lastOff, c.dTraceOff.p + (n - 1), 8, cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&lastWords, c.dPlaneWordsB.p + (n - 1), 8, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    if (st)
        st->kernel_launches += 2 + 2;
    unsigned int classStart[NC + 1];
    classStart[0] = 0;
    for (int cls = 0; cls < NC; ++cls)
        classStart[cls + 1] = classStart[cls] + info[cls];
    unsigned int const nPacked = classStart[NP];
    uint64_t const     total   = lastOff + lastWords;

    // groups of sorted slots whose planes fit the budget of one launch (almost always one group)
    uint64_t const kMaxPlaneWords = c.maxPlaneWords;
    struct Group
    {
        unsigned int b, e;
        uint64_t     base, words;
    };
    std::vector<Group> groups;
    if (nPacked)
    {
        if (total <= kMaxPlaneWords)
            groups.push_back({0u, nPacked, 0ull, total});
        else
        {
            std::vector<unsigned long long> off(nPacked + 1);
            LGPU_CUDA(cudaMemcpyAsync(off.data(), c.dTraceOff.p, nPacked * 8ull, cudaMemcpyDeviceToHost, c.stream));
            syncStream(c);
            off[nPacked] = total;
            unsigned int b = 0;
            while (b < nPacked)
            {
                unsigned int e = b + 1;
                while (e < nPacked && off[e + 1] - off[b] <= kMaxPlaneWords)
                    ++e;
                groups.push_back({b, e, off[b], off[e] - off[b]});
                b = e;
            }
        }
    }
    unsigned int nNonEmpty = 0;
    for (int cls = 0; cls < NP; ++cls)
        nNonEmpty += info[cls] ? 1u : 0u;
    for (Group const & gr : groups)
    {
        c.dPlanes.reserve(gr.words);
        unsigned int kLaunch = 0;
        for (int cls = 0; cls < NP; ++cls)
        {
            unsigned int const lo = std::max(classStart[cls], gr.b), hi = std::min(classStart[cls + 1], gr.e);
            if (lo >= hi)
                continue;
            unsigned int const G = dpxGroupsOf(tab, cls);
            // one launch group: the host has just waited for the stream, so the classes may fan out over side streams;
            // several groups reuse the plane buffer one after the other and stay on the context's stream
            cudaStream_t const cs = groups.size() == 1 ? classStream(c, kLaunch++, nNonEmpty) : c.stream;
            LGPU_CUDA(cudaMemsetAsync(c.dWork.p + cls, 0, 4, cs));
            DpxParams P    = baseDpxParams(c, dTasks, n);
            P.order        = c.dOrderB.p;
            P.keys         = c.dClassKeysB.p;
            P.nJobs        = (hi - lo + G - 1) / G;
            P.slotBase     = lo;
            P.nSlots       = hi - lo;
            P.workCounter  = c.dWork.p + cls;
            P.scores       = c.dScores2.p;
            P.planes       = c.dPlanes.p;
            P.planeOff     = c.dTraceOff.p;
            P.planeOffBase = gr.base;
            P.bestCol      = c.dBestPos.p;
            (tab == kDpxTabPriv ? kDpxLaunchPrivTr[cls] : kDpxLaunchTrace32[cls])(c, P, info[NC + cls], cs);
            if (st)
                st->kernel_launches += 1;
        }
        joinClassStreams(c);
        TracebackResParams TP{};
        TP.ix           = c.index->dev;
        TP.Q            = c.Q;
        TP.tasks        = dTasks;
        TP.order        = c.dOrderB.p;
        TP.keys         = c.dClassKeysB.p;
        TP.slotFirst    = gr.b;
        TP.nSlots       = gr.e - gr.b;
        TP.tab          = tab;
        TP.matrix       = c.dMatrix.p;
        TP.go           = c.scoring.gapOpenSeqan;
        TP.ge           = c.scoring.gapExtend;
        TP.scores       = c.dScores2.p;
        TP.bestCol      = c.dBestPos.p;
        TP.planes       = c.dPlanes.p;
        TP.planeOff     = c.dTraceOff.p;
        TP.planeOffBase = gr.base;
        TP.out          = c.dHits.p;
        unsigned int const tbGrid = gridFor(static_cast<unsigned long long>(TP.nSlots) * 32, 32 * kTbWarps);
        tracebackResKernel<<<tbGrid, 32 * kTbWarps, 0, c.stream>>>(TP);
        LGPU_CUDA(cudaGetLastError());
        if (st)
            st->kernel_launches += 1;
        if (c.params.want_cigar)
            emitCigars(c, c.dOrderB.p + gr.b, TP.nSlots, [&](unsigned int * ops, unsigned int const * offs, unsigned int base) {
                TP.emit      = 1;
                TP.cigarOps  = ops;
                TP.cigarOff  = offs;
                TP.cigarBase = base;
                tracebackResKernel<<<tbGrid, 32 * kTbWarps, 0, c.stream>>>(TP);
            }, st);
    }

    // ---- scalar class: the last slots of the sorted order ----
    unsigned int const nScalar = n - nPacked;
    if (nScalar)
    {
        c.dTasksScalar.reserve(nScalar);
        gatherTasksKernel<<<gridFor(nScalar, 256), 256, 0, c.stream>>>(dTasks, c.dOrderB.p + nPacked, nScalar, c.dTasksScalar.p);
        std::vector<lgpu_match> sub(nScalar);
        LGPU_CUDA(cudaMemcpyAsync(sub.data(), c.dTasksScalar.p, nScalar * sizeof(lgpu_match), cudaMemcpyDeviceToHost, c.stream));
        syncStream(c);
        // the sorted order is not needed by the scalar kernels' scratch arrays, but its tail is the output index
        c.dScalarIdx.reserve(nScalar);
        LGPU_CUDA(cudaMemcpyAsync(c.dScalarIdx.p, c.dOrderB.p + nPacked, nScalar * 4ull, cudaMemcpyDeviceToDevice, c.stream));
        runTraceScalar(c, sub.data(), nScalar, c.dTasksScalar.p, c.dScalarIdx.p, st);
        if (st)
            st->kernel_launches += 1;
    }
    if (st)
    {
        st->n_extensions_trace += n;
        st->cells_trace += cells;
    }
}

// iterateMatches for one phase: c.dMatches[0..nMatches) -> records appended to c.dAllHits (device)
static void runExtension(lgpu_ctx & c, uint64_t nMatches, uint8_t phase, lgpu_stats * st)
{
    uint64_t const nTasks = runMerge(c, c.dMatches.p, nMatches, st);
    if (nTasks == 0)
        return;
    c.dScores.reserve(nTasks);
    traceTime("extension: merged");
    runScorePass(c, c.dMerged.p, static_cast<unsigned int>(nTasks), c.dScores.p, st);
    traceTime("extension: score pass launched");

    // filter on the device with per-query integer thresholds
    c.dHead.reserve(nTasks);
    c.dScan.reserve(nTasks);
    c.dTasks2.reserve(nTasks);
    LGPU_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 8 * sizeof(unsigned long long), c.stream));
    unsigned int const g = gridFor(nTasks, 256);
    filterKernel<<<g, 256, 0, c.stream>>>(c.dMerged.p, c.dScores.p, static_cast<unsigned int>(nTasks), c.Q.F, c.dMinBit.p,
                                          c.dMinEval.p, c.dHead.p, c.dCounters.p);
    size_t tmp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp, c.dHead.p, c.dScan.p, static_cast<int>(nTasks), c.stream);
    c.dCubTemp.reserve(tmp);
    size_t tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::InclusiveSum(c.dCubTemp.p, tb, c.dHead.p, c.dScan.p, static_cast<int>(nTasks), c.stream));
    compactKernel<<<g, 256, 0, c.stream>>>(c.dMerged.p, c.dHead.p, c.dScan.p, static_cast<unsigned int>(nTasks), c.dTasks2.p);
    LGPU_CUDA(cudaGetLastError());
    unsigned int       nKeep = 0;
    unsigned long long cnt[2];
    LGPU_CUDA(cudaMemcpyAsync(&nKeep, c.dScan.p + (nTasks - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(cnt, c.dCounters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    if (st)
    {
        st->kernel_launches += 3;
        st->hits_failed_bitscore += cnt[0];
        st->hits_failed_evalue += cnt[1];
    }
    if (nKeep == 0)
        return;
    traceTime("extension: filtered");
    runTracePass(c, c.dTasks2.p, nKeep, st);
    traceTime("extension: trace pass launched");

    // identity cut-off, phase, "query is done" flags (src/search_algo.hpp:1308-1322); survivors join the batch's records
    c.dAllHits.grow(c.nAll + nKeep, c.nAll, c.stream);
    c.dHead.reserve(nKeep);
    c.dScan.reserve(nKeep);
    LGPU_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 8, c.stream));
    unsigned int const g2 = gridFor(nKeep, 256);
    postTraceKernel<<<g2, 256, 0, c.stream>>>(c.dHits.p, nKeep, c.params.id_cutoff, phase, c.dHead.p, c.dQryHasHit.p, c.dCounters.p);
    if (st)
        st->kernel_launches += 1;
    if (c.params.id_cutoff > 0)
    {
        tb = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb, c.dHead.p, c.dScan.p, static_cast<int>(nKeep), c.stream);
        c.dCubTemp.reserve(tb);
        tb = c.dCubTemp.cap;
        LGPU_CUDA(cub::DeviceScan::InclusiveSum(c.dCubTemp.p, tb, c.dHead.p, c.dScan.p, static_cast<int>(nKeep), c.stream));
        appendHitsKernel<<<g2, 256, 0, c.stream>>>(c.dHits.p, c.dHead.p, c.dScan.p, nKeep, c.dAllHits.p + c.nAll);
        unsigned int       nOk = 0;
        unsigned long long nFailed = 0;
        LGPU_CUDA(cudaMemcpyAsync(&nOk, c.dScan.p + (nKeep - 1), 4, cudaMemcpyDeviceToHost, c.stream));
        LGPU_CUDA(cudaMemcpyAsync(&nFailed, c.dCounters.p, 8, cudaMemcpyDeviceToHost, c.stream));
        syncStream(c);
        c.nAll += nOk;
        if (st)
        {
            st->kernel_launches += 2;
            st->hits_failed_identity += nFailed;
        }
    }
    else
    {
        // an identity below 0 does not exist: every record survives, in task order
        LGPU_CUDA(cudaMemcpyAsync(c.dAllHits.p + c.nAll, c.dHits.p, nKeep * sizeof(lgpu_hit), cudaMemcpyDeviceToDevice, c.stream));
        c.nAll += nKeep;
    }
}

// per-query integer thresholds of the pass-1 filter; the per-length values are cached across batches
static void setThresholds(lgpu_ctx & c, lgpu_stats * st)
{
    HostTimer            ht(st ? &st->ms_host : nullptr);
    uint64_t const       n = c.nQueries;
    constexpr uint64_t   kDirect = 1u << 16;
    if (c.thrByLen.empty())
    {
        c.thrByLen.resize(kDirect);
        c.thrKnown.assign(kDirect, 0);
    }
    c.hMinBit.reserve(n);
    c.hMinEval.reserve(n);
    for (uint64_t q = 0; q < n; ++q)
    {
        uint64_t const  len = c.qOffsHost[q + 1] - c.qOffsHost[q];
        ScoreThresholds t;
        if (len < kDirect)
        {
            if (!c.thrKnown[len])
            {
                c.thrByLen[len] = scoreThresholds(c.params, *c.evc, len);
                c.thrKnown[len] = 1;
            }
            t = c.thrByLen[len];
        }
        else
        {
            auto it = c.thrLong.find(len);
            if (it == c.thrLong.end())
                it = c.thrLong.emplace(len, scoreThresholds(c.params, *c.evc, len)).first;
            t = it->second;
        }
        c.hMinBit.p[q]  = t.minBit;
        c.hMinEval.p[q] = t.minEval;
    }
    c.dMinBit.reserve(n);
    c.dMinEval.reserve(n);
    LGPU_CUDA(cudaMemcpyAsync(c.dMinBit.p, c.hMinBit.p, n * 4, cudaMemcpyHostToDevice, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(c.dMinEval.p, c.hMinEval.p, n * 4, cudaMemcpyHostToDevice, c.stream));
}

// iterativeSearchPre/Post (src/search_algo.hpp:1391-1460): the queries without a reported hit after phase 1 form
// the active list of phase 2; built on the device, the host learns its size and the longest active query
static unsigned int buildActiveList(lgpu_ctx & c, unsigned int & maxLen, lgpu_stats * st)
{
    unsigned int const n = static_cast<unsigned int>(c.nQueries);
    unsigned int const g = gridFor(n, 256);
    c.dHead.reserve(n);
    c.dScan.reserve(n);
    c.dClassInfo.reserve(4);
    LGPU_CUDA(cudaMemsetAsync(c.dClassInfo.p, 0, 4, c.stream));
    notDoneKernel<<<g, 256, 0, c.stream>>>(c.dQryHasHit.p, n, c.dHead.p);
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, c.dHead.p, c.dScan.p, static_cast<int>(n), c.stream);
    c.dCubTemp.reserve(tb);
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::InclusiveSum(c.dCubTemp.p, tb, c.dHead.p, c.dScan.p, static_cast<int>(n), c.stream));
    activeEmitKernel<<<g, 256, 0, c.stream>>>(c.dHead.p, c.dScan.p, c.dQOffs.p, n, c.dActive.p, c.dClassInfo.p);
    LGPU_CUDA(cudaGetLastError());
    unsigned int nActive = 0;
    maxLen               = 0;
    LGPU_CUDA(cudaMemcpyAsync(&nActive, c.dScan.p + (n - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&maxLen, c.dClassInfo.p, 4, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    if (st)
        st->kernel_launches += 3;
    return nActive;
}

// writeRecords / _writeRecord on the device (kernels_finalize.cuh): c.dAllHits[0..nAll) -> c.finalDev[0..nFinal)
static void finalizeOnDevice(lgpu_ctx & c, lgpu_stats * st)
{
    c.finalDev = c.dAllHits.p;
    c.nFinal   = c.nAll;
    if (!c.params.finalize || c.nAll == 0)
        return;
    if (c.nAll >= (1ull << 31))
        throw ArgError("too many records in one batch; use smaller query batches");
    unsigned int const n  = static_cast<unsigned int>(c.nAll);
    int const          nI = static_cast<int>(n);
    unsigned int const g  = gridFor(n, 256);
    c.dFinIdx.reserve(n);
    c.dFinIdx1.reserve(n);
    c.dDropped.reserve(n);
    c.dSegStart.reserve(n);
    c.dSegStartB.reserve(n);
    c.dHead.reserve(n);
    c.dScan.reserve(n);
    c.dFinal.reserve(n);
    LGPU_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 8 * sizeof(unsigned long long), c.stream));
    iotaKernel<<<g, 256, 0, c.stream>>>(c.dFinIdx.p, n);
    RecordLess const less1{c.dAllHits.p};
    RankLess const   less2{c.dAllHits.p, c.dDropped.p};
    size_t           t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    cub::DeviceMergeSort::StableSortKeys(nullptr, t1, c.dFinIdx.p, nI, less1, c.stream);
    cub::DeviceMergeSort::StableSortKeys(nullptr, t2, c.dFinIdx.p, nI, less2, c.stream);
    cub::DeviceScan::InclusiveScan(nullptr, t3, c.dSegStart.p, c.dSegStartB.p, MaxOp(), nI, c.stream);
    cub::DeviceScan::InclusiveSum(nullptr, t4, c.dHead.p, c.dScan.p, nI, c.stream);
    c.dCubTemp.reserve(std::max(std::max(t1, t2), std::max(t3, t4)));
    size_t tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceMergeSort::StableSortKeys(c.dCubTemp.p, tb, c.dFinIdx.p, nI, less1, c.stream));
    markDuplicatesKernel<<<g, 256, 0, c.stream>>>(c.dAllHits.p, c.dFinIdx.p, n, c.dDropped.p, c.dCounters.p);
    LGPU_CUDA(cudaMemcpyAsync(c.dFinIdx1.p, c.dFinIdx.p, n * 4ull, cudaMemcpyDeviceToDevice, c.stream));
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceMergeSort::StableSortKeys(c.dCubTemp.p, tb, c.dFinIdx.p, nI, less2, c.stream));
    queryHeadKernel<<<g, 256, 0, c.stream>>>(c.dAllHits.p, c.dFinIdx.p, c.dDropped.p, n, c.dSegStart.p);
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::InclusiveScan(c.dCubTemp.p, tb, c.dSegStart.p, c.dSegStartB.p, MaxOp(), nI, c.stream));
    rankCutKernel<<<g, 256, 0, c.stream>>>(c.dFinIdx.p, c.dDropped.p, c.dSegStartB.p, n, c.params.max_matches, c.dHead.p,
                                           c.dCounters.p + 1);
    tb = c.dCubTemp.cap;
    LGPU_CUDA(cub::DeviceScan::InclusiveSum(c.dCubTemp.p, tb, c.dHead.p, c.dScan.p, nI, c.stream));
    gatherFinalKernel<<<g, 256, 0, c.stream>>>(c.dAllHits.p, c.dFinIdx.p, c.dHead.p, c.dScan.p, n, c.dFinal.p);
    countPairsKernel<<<g, 256, 0, c.stream>>>(c.dAllHits.p, c.dFinIdx1.p, c.dDropped.p, n, c.dCounters.p + 3);
    LGPU_CUDA(cudaGetLastError());
    unsigned long long cnt[4];
    unsigned int       nFin = 0;
    LGPU_CUDA(cudaMemcpyAsync(cnt, c.dCounters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c.stream));
    LGPU_CUDA(cudaMemcpyAsync(&nFin, c.dScan.p + (n - 1), 4, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    c.finalDev = c.dFinal.p;
    c.nFinal   = nFin;
    if (st)
    {
        st->kernel_launches += 7 + 4;
        st->hits_duplicate2 += cnt[0];
        st->hits_abundant += cnt[1];
        st->qrys_with_hit += cnt[2];
        st->pairs += cnt[3];
        st->hits_final += nFin;
    }
}

// the device part of the path for one batch on one context / one stream; the records end up in c.finalDev[0..nFinal)
static void searchDevice(lgpu_ctx & c, BatchView const & qb, lgpu_stats * st)
{
    LGPU_CUDA(cudaSetDevice(c.index->device));
    SearchThreadScope const scope; // counted by the host-wait policy (effectiveSyncMode)
    c.cigar.clear();
    c.nAll       = 0;
    c.nFinal     = 0;
    c.finalDev   = nullptr;
    c.timersUsed = 0;
    traceTime("search: start");
    uploadQueries(c, qb, st);
    if (!c.evc)
        c.evc = std::make_unique<EValueComputer>(c.scoring.ka, c.index->dbTotalLength, c.di.qIsTranslated);
    if (!c.nQueries)
        return;
    setThresholds(c, st);
    unsigned int const n = static_cast<unsigned int>(c.nQueries);
    c.dActive.reserve(n);
    c.dQryHasHit.reserve(n);
    iotaKernel<<<gridFor(n, 256), 256, 0, c.stream>>>(c.dActive.p, n);
    LGPU_CUDA(cudaMemsetAsync(c.dQryHasHit.p, 0, n * 4ull, c.stream));
    if (st)
        st->kernel_launches += 1;
    if (c.params.iterative_search)
    {
        uint64_t nM = runSeeding(c, c.params.opts0, c.dActive.p, n, c.maxQueryLen, st);
        traceTime("search: phase 1 seeded");
        runExtension(c, nM, 1, st);
        traceTime("search: phase 1 extended");
        unsigned int       maxLen  = 0;
        unsigned int const nActive = buildActiveList(c, maxLen, st);
        if (nActive)
        {
            nM = runSeeding(c, c.params.opts, c.dActive.p, nActive, maxLen, st);
            runExtension(c, nM, 2, st);
        }
    }
    else
    {
        uint64_t const nM = runSeeding(c, c.params.opts, c.dActive.p, n, c.maxQueryLen, st);
        runExtension(c, nM, 2, st);
    }
    traceTime("search: phases done");
    finalizeOnDevice(c, st);
    traceTime("search: finalised");
}

// Second half of a search: the nFinal records of searchDevice() to `dst` (pinned host memory of the calling context),
// query ids / cigar offsets rebased for sub-batches; the doubles (bit score, e-value) come from the host's tables.
static void fetchRecords(lgpu_ctx & c, lgpu_hit * dst, uint32_t qBase, uint32_t cigarBase, lgpu_stats * st)
{
    LGPU_CUDA(cudaSetDevice(c.index->device));
    if (c.nFinal)
    {
        StageTimer t(c, st ? &st->ms_d2h : nullptr);
        LGPU_CUDA(cudaMemcpyAsync(dst, c.finalDev, c.nFinal * sizeof(lgpu_hit), cudaMemcpyDeviceToHost, c.stream));
    }
    LGPU_CUDA(cudaMemcpyAsync(c.hOverflow.p, c.dOverflow.p, 4, cudaMemcpyDeviceToHost, c.stream));
    syncStream(c);
    if (c.hOverflow.p[0])
        checkScoreOverflow(c);
    if (c.nFinal)
    {
        HostTimer ht(st ? &st->ms_host : nullptr);
        for (uint64_t i = 0; i < c.nFinal; ++i)
        {
            lgpu_hit & h = dst[i];
            h.bit_score  = c.evc->bitsCached(h.score);
            h.evalue     = c.evc->evalueCached(h.score, h.q_len);
            h.q_id += qBase;
            h.cigar_off += cigarBase;
        }
    }
    resolveTimers(c);
    traceTime("search: records on the host");
}

static std::unique_ptr<lgpu_ctx> makeContext(lgpu_index const * ix, lgpu_params const & p);

static void addStats(lgpu_stats & dst, lgpu_stats const & src)
{
    uint64_t *       d = reinterpret_cast<uint64_t *>(&dst);
    uint64_t const * s = reinterpret_cast<uint64_t const *>(&src);
    for (int k = 0; k < 16; ++k)
        d[k] += s[k];
    dst.ms_seed += src.ms_seed;
    dst.ms_sort_merge += src.ms_sort_merge;
    dst.ms_extend_score += src.ms_extend_score;
    dst.ms_extend_trace += src.ms_extend_trace;
    dst.ms_h2d += src.ms_h2d;
    dst.ms_d2h += src.ms_d2h;
    dst.ms_host += src.ms_host;
}

// Host offsets of a public batch (D2H when the caller's arrays live on the device).
static BatchView viewOf(lgpu_ctx & c, lgpu_query_batch const & qb, std::vector<uint64_t> & offs)
{
    offs.resize(qb.n_queries + 1);
    if (qb.n_queries)
    {
        if (qb.on_device)
        {
            LGPU_CUDA(cudaMemcpyAsync(offs.data(), qb.offsets, (qb.n_queries + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
            syncStream(c);
        }
        else
            std::memcpy(offs.data(), qb.offsets, (qb.n_queries + 1) * 8);
    }
    else
        offs[0] = 0;
    BatchView v;
    v.residues    = qb.residues;
    v.resOnDevice = qb.on_device != 0;
    v.offsets     = offs.data();
    v.n           = qb.n_queries;
    return v;
}

// One public call.  Large batches are cut into `streams` contiguous sub-batches that run the whole
// pipeline concurrently on their own CUDA stream and host thread: the host-only and latency-bound
// stretches of one sub-batch (score thresholds, record finalisation, the phase-2 stragglers, every
// device->host hand-over) overlap with the DP kernels of the other.  Results do not depend on the
// split (every query is independent).
static void searchBatch(lgpu_ctx & c, lgpu_query_batch const & qb, lgpu_hits * out, lgpu_stats * st)
{
    LGPU_CUDA(cudaSetDevice(c.index->device));
    std::vector<uint64_t> offs;
    BatchView const       all = viewOf(c, qb, offs);
    LGPU_CUDA(cudaEventRecord(c.ev[2], c.stream));
    // sub-batches below ~16k queries do not fill the GPU any more (measured: 10k real-length queries run
    // 24.5 ms on one stream, 29.9 ms cut in three); tests force the split with LAMBDA_B200_MIN_SUBBATCH
    uint64_t const     minSub = c.minSubBatch;
    unsigned int const nW = (c.streams > 1 && all.n >= 2 * minSub)
                              ? static_cast<unsigned int>(std::min<uint64_t>(c.streams, all.n / minSub))
                              : 1;
    c.lastParts.clear();
    c.nResult = 0;
    if (nW == 1)
    {
        searchDevice(c, all, st);
        c.hResult.reserve(c.nFinal);
        fetchRecords(c, c.hResult.p, 0, 0, st);
        c.nResult = c.nFinal;
        c.lastParts.push_back({&c, 0ull, 0ull, c.nFinal});
    }
    else
    {
        while (c.workers.size() < nW)
        {
            c.workers.push_back(makeContext(c.index, c.params));
            // Sub-batches share the GPU: the ALU-bound DP score kernel leaves shared memory and warp slots
            // for the memory-bound stages (seeding, trace fill, traceback) of the other sub-batches
            // (measured: profiles/r1_sweep_streams.jsonl).
            if (!std::getenv("LAMBDA_B200_DPX_OCC"))
                c.workers.back()->dpxBlocksPerSM = 10;
        }
        std::vector<std::vector<uint64_t>> subOffs(nW);
        std::vector<lgpu_stats>            wst(nW);
        std::vector<std::string>           err(nW);
        std::memset(wst.data(), 0, sizeof(lgpu_stats) * nW);
        // every worker thread runs stage 1 (device) of its sub-batch, waits until all record counts are known, then
        // stage 2: its records straight into its slice of the call's result buffer
        struct Gate
        {
            std::mutex              m;
            std::condition_variable cv;
            unsigned int            arrived = 0;
            bool                    open    = false;
        } gate;
        std::vector<uint64_t> recOff(nW + 1, 0), cigOff(nW + 1, 0);
        std::vector<std::thread> th;
        struct Joiner // an exception while spawning must not destroy joinable threads
        {
            std::vector<std::thread> & t;
            ~Joiner()
            {
                for (auto & x : t)
                    if (x.joinable())
                        x.join();
            }
        } joiner{th};
        std::vector<BatchView> views(nW);
        for (unsigned int w = 0; w < nW; ++w) // everything that may fail before a thread exists
        {
            uint64_t const b = all.n * w / nW, e = all.n * (w + 1) / nW;
            subOffs[w].assign(all.offsets + b, all.offsets + e + 1);
            uint64_t const base = subOffs[w][0];
            for (auto & x : subOffs[w])
                x -= base;
            views[w].residues    = all.residues + base;
            views[w].resOnDevice = all.resOnDevice;
            views[w].offsets     = subOffs[w].data();
            views[w].n           = e - b;
            LGPU_CUDA(cudaStreamWaitEvent(c.workers[w]->stream, c.ev[2], 0));
        }
        th.reserve(nW);
        for (unsigned int w = 0; w < nW; ++w)
        {
            uint64_t const  b  = all.n * w / nW;
            BatchView const v  = views[w];
            lgpu_ctx *      wc = c.workers[w].get();
            try
            {
            th.emplace_back([&, wc, v, w, b, nW] {
                try
                {
                    searchDevice(*wc, v, &wst[w]);
                }
                catch (std::exception const & ex)
                {
                    err[w]     = ex.what();
                    wc->nFinal = 0;
                }
                {
                    // the last thread to arrive lays out the result buffer
                    std::unique_lock<std::mutex> lock(gate.m);
                    if (++gate.arrived >= nW)
                    {
                        try
                        {
                            for (unsigned int k = 0; k < nW; ++k)
                            {
                                recOff[k + 1] = recOff[k] + c.workers[k]->nFinal;
                                cigOff[k + 1] = cigOff[k] + c.workers[k]->cigar.size();
                            }
                            if (cigOff[nW] > 0xffffffffull)
                                throw ArgError("too many alignment operations in one batch; use smaller batches with want_cigar");
                            cudaSetDevice(c.index->device);
                            c.hResult.reserve(recOff[nW]);
                        }
                        catch (std::exception const & ex)
                        {
                            err[w] = ex.what();
                        }
                        gate.open = true;
                        gate.cv.notify_all();
                    }
                    else
                        gate.cv.wait(lock, [&] { return gate.open; });
                }
                bool anyErr = false;
                for (auto const & e2 : err)
                    anyErr = anyErr || !e2.empty();
                if (anyErr)
                    return;
                try
                {
                    fetchRecords(*wc, c.hResult.p + recOff[w], static_cast<uint32_t>(b), static_cast<uint32_t>(cigOff[w]), &wst[w]);
                    LGPU_CUDA(cudaEventRecord(wc->ev[4], wc->stream));
                }
                catch (std::exception const & ex)
                {
                    err[w] = ex.what();
                }
            });
            }
            catch (std::exception const & ex) // the thread could not be started: let the others through the gate
            {
                std::unique_lock<std::mutex> lock(gate.m);
                for (unsigned int k = w; k < nW; ++k)
                    err[k] = std::string("could not start a worker thread: ") + ex.what();
                gate.arrived += nW - w;
                if (gate.arrived >= nW)
                {
                    gate.open = true;
                    gate.cv.notify_all();
                }
                break;
            }
        }
        for (auto & t : th)
            t.join();
        for (unsigned int w = 0; w < nW; ++w)
            if (!err[w].empty())
                throw CudaError("worker " + std::to_string(w) + ": " + err[w]);
        c.cigar.clear();
        for (unsigned int w = 0; w < nW; ++w)
        {
            lgpu_ctx & wc = *c.workers[w];
            c.lastParts.push_back({&wc, all.n * w / nW, cigOff[w], wc.nFinal});
            c.cigar.insert(c.cigar.end(), wc.cigar.begin(), wc.cigar.end());
            LGPU_CUDA(cudaStreamWaitEvent(c.stream, wc.ev[4], 0));
            if (st)
                addStats(*st, wst[w]);
        }
        c.nResult = recOff[nW];
    }
    LGPU_CUDA(cudaEventRecord(c.ev[3], c.stream));
    waitEvent(c, c.ev[3]);
    if (st)
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]);
        st->ms_total += ms;
    }
    out->hits        = c.hResult.p;
    out->n           = c.nResult;
    out->cigar_ops   = c.params.want_cigar ? c.cigar.data() : nullptr;
    out->n_cigar_ops = c.params.want_cigar ? c.cigar.size() : 0;
}

} // namespace lgpu

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------

template <typename F>
static int guarded(std::string * err, F && f)
{
    try
    {
        f();
        return LGPU_OK;
    }
    catch (LbaError const & e)
    {
        (err ? *err : g_lastError) = e.what();
        g_lastError                = e.what();
        return LGPU_ERR_IO;
    }
    catch (CudaError const & e)
    {
        (err ? *err : g_lastError) = e.what();
        g_lastError                = e.what();
        return LGPU_ERR_CUDA;
    }
    catch (ArgError const & e)
    {
        (err ? *err : g_lastError) = e.what();
        g_lastError                = e.what();
        return LGPU_ERR_ARG;
    }
    catch (UnsupportedError const & e)
    {
        (err ? *err : g_lastError) = e.what();
        g_lastError                = e.what();
        return LGPU_ERR_UNSUPPORTED;
    }
    catch (std::exception const & e)
    {
        (err ? *err : g_lastError) = e.what();
        g_lastError                = e.what();
        return LGPU_ERR_INTERNAL;
    }
}

static std::unique_ptr<lgpu_ctx> lgpu::makeContext(lgpu_index const * ix, lgpu_params const & p)
{
    checkParams(p, ix->meta);
    auto c    = std::make_unique<lgpu_ctx>();
    c->index  = ix;
    c->device = ix->device;
    c->params = p;
    c->di     = domainInfo(p.domain, ix->meta.orig_alph, p.query_alph);
    int rc    = makeScoring(c->scoring, p);
    if (rc == LGPU_ERR_ARG)
        throw ArgError("Could not compute Karlin-Altschul-Values for Scoring Scheme.");
    if (rc != LGPU_OK)
        throw UnsupportedError("unsupported scoring configuration");
    LGPU_CUDA(cudaSetDevice(ix->device));
    LGPU_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (char const * e = std::getenv("LAMBDA_B200_BLOCKING_SYNC"))
        c->syncMode = std::atoi(e) != 0 ? LGPU_SYNC_BLOCK : LGPU_SYNC_AUTO;
    if (char const * e = std::getenv("LAMBDA_B200_SYNC"))
        c->syncMode = !std::strcmp(e, "spin") ? LGPU_SYNC_SPIN : !std::strcmp(e, "block") ? LGPU_SYNC_BLOCK
                      : !std::strcmp(e, "yield") ? LGPU_SYNC_YIELD : LGPU_SYNC_AUTO;
    for (auto & e : c->ev) // stage timers: waited on by the host as well
        LGPU_CUDA(cudaEventCreateWithFlags(&e, c->syncMode == LGPU_SYNC_BLOCK ? cudaEventBlockingSync : cudaEventDefault));
    LGPU_CUDA(cudaEventCreateWithFlags(&c->evSync, cudaEventBlockingSync | cudaEventDisableTiming));
    LGPU_CUDA(cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, ix->device));
    c->dMatrix.reserve(2048);
    LGPU_CUDA(cudaMemcpy(c->dMatrix.p, c->scoring.matrix, 1024, cudaMemcpyHostToDevice));
    LGPU_CUDA(cudaMemcpy(c->dMatrix.p + 1024, c->scoring.matrixRev, 1024, cudaMemcpyHostToDevice));
    c->dCounters.reserve(8);
    c->dOverflow.reserve(1);
    c->hOverflow.reserve(1);
    LGPU_CUDA(cudaMemset(c->dOverflow.p, 0, 4));
    if (char const * e = std::getenv("LAMBDA_B200_SEED"))
        c->seedMode = !std::strcmp(e, "thread") ? 1 : !std::strcmp(e, "warp") ? 2 : !std::strcmp(e, "block") ? 3 : !std::strcmp(e, "spec") ? 4 : 0;
    if (char const * e = std::getenv("LAMBDA_B200_CLASS_STREAMS"))
        c->classStreams = std::atoi(e) != 0;
    if (char const * e = std::getenv("LAMBDA_B200_SEED_TEXT"))
        c->seedTextElong = std::atoi(e) != 0;
    if (char const * e = std::getenv("LAMBDA_B200_SEED_PREFIX"))
        c->seedPrefix = std::atoi(e) != 0;
    if (char const * e = std::getenv("LAMBDA_B200_DPX_OCC"))
        c->dpxBlocksPerSM = static_cast<unsigned int>(std::max(1, std::min(32, std::atoi(e))));
    if (char const * e = std::getenv("LAMBDA_B200_TRACE"))
    {
        c->forceScalarTrace = !std::strcmp(e, "scalar");
    }
    // the packed kernel stores (score - gapOpen) as int8 profile bytes with -128 reserved for "null"
    c->dpxOk = c->scoring.alphSize < 32 && c->scoring.gapOpenSeqan <= c->scoring.gapExtend && c->scoring.gapExtend <= 0;
    for (int a = 0; a < c->scoring.alphSize; ++a)
        for (int b = 0; b < c->scoring.alphSize; ++b)
        {
            int const v = c->scoring.matrix[a * 32 + b] - c->scoring.gapOpenSeqan;
            int const w = c->scoring.matrixRev[a * 32 + b] - c->scoring.gapOpenSeqan;
            if (v < -127 || v > 127 || w < -127 || w > 127)
                c->dpxOk = false;
        }
    c->dpxScoreOk = c->dpxOk;
    // Small alphabets (nucleotide / bisulfite searches: about one alignment per read) build one query profile per
    // group instead of one per warp, so the groups of a warp can carry different reads.  LAMBDA_B200_PROFILE=
    // shared|private forces either (tests).
    c->privProfiles = c->scoring.alphSize + 1 <= 8;
    if (char const * e = std::getenv("LAMBDA_B200_PROFILE"))
        c->privProfiles = !std::strcmp(e, "private") ? true : !std::strcmp(e, "shared") ? false : c->privProfiles;
    c->traceTab = c->privProfiles ? kDpxTabPriv : kDpxTabTrace32;
    // residue-plane traceback: neighbouring cells must differ by less than half the residue range
    {
        int mx = -128, mn = 127;
        for (int a = 0; a < c->scoring.alphSize; ++a)
            for (int b = 0; b < c->scoring.alphSize; ++b)
            {
                mx = std::max(mx, std::max<int>(c->scoring.matrix[a * 32 + b], c->scoring.matrixRev[a * 32 + b]));
                mn = std::min(mn, std::min<int>(c->scoring.matrix[a * 32 + b], c->scoring.matrixRev[a * 32 + b]));
            }
        c->maxColsInt16 = static_cast<unsigned int>(32767 / std::max(mx, 1));
        int const go    = c->scoring.gapOpenSeqan;
        c->resTraceOk = c->dpxOk && go < 0 && c->scoring.gapExtend <= 0 && mx - go <= 127 && 2 * (mx - go) - mn < 256;
    }
    if (char const * e = std::getenv("LAMBDA_B200_PLANE_MB")) // tests: force several launch groups
        c->maxPlaneWords = static_cast<uint64_t>(std::max(1, std::atoi(e))) * (1u << 20) / 4;
    traceTime("context created");
    return c;
}

extern "C"
{

int lgpu_version(void) { return LGPU_VERSION; }

int lgpu_lba_open(lgpu_lba ** out, char const * path)
{
    if (!out || !path)
        return LGPU_ERR_ARG;
    *out = nullptr;
    traceTime("lba: open");
    return guarded(nullptr, [&] {
        auto l  = std::make_unique<lgpu_lba>();
        l->file = std::make_unique<LbaFile>(path);
        *out    = l.release();
    });
}

lgpu_index_desc const * lgpu_lba_desc(lgpu_lba const * l) { return l ? &l->file->desc : nullptr; }

lgpu_taxonomy const * lgpu_lba_taxonomy(lgpu_lba const * l) { return l ? &l->file->tax : nullptr; }

void lgpu_lba_close(lgpu_lba * l) { delete l; }

int lgpu_index_create(lgpu_index ** out, lgpu_index_desc const * d, int device)
{
    if (!out || !d)
        return LGPU_ERR_ARG;
    *out = nullptr;
    traceTime("index: create called");
    return guarded(nullptr, [&] {
        if (d->index_type != LGPU_INDEX_FM)
            throw UnsupportedError("only unidirectional FM indexes are supported");
        if (d->sigma < 2 || d->sigma > 31 || d->sigma_bits > 5)
            throw ArgError("bad alphabet size in index descriptor");
        int nDev = 0;
        if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
            throw CudaError("no CUDA device available (lambda_b200 has no CPU fallback)");
        LGPU_CUDA(cudaSetDevice(device));
        traceTime("index: context ready");
        auto ix    = std::make_unique<lgpu_index>();
        ix->device = device;
        std::vector<UploadJob> pending; // large blobs are copied together by runUploads()
        auto up    = [&](void const * p, uint64_t bytes) -> void * {
            void * dp = nullptr;
            if (bytes < (8u << 20))
                dp = uploadNew(static_cast<unsigned char const *>(p), bytes);
            else
            {
                LGPU_CUDA(cudaMalloc(&dp, bytes));
                pending.push_back({static_cast<unsigned char *>(dp), static_cast<unsigned char const *>(p), bytes});
            }
            ix->allocs.push_back(dp);
            ix->allocBytes.push_back(static_cast<size_t>(bytes));
            ix->bytes += bytes;
            return dp;
        };
        DevIndex & dv  = ix->dev;
        dv.occ         = static_cast<unsigned char const *>(up(d->occ_blocks, d->n_blocks * d->block_bytes));
        dv.super       = static_cast<unsigned long long const *>(up(d->super_blocks, d->n_super * d->sigma * 8));
        dv.ssa         = static_cast<unsigned long long const *>(up(d->ssa, d->n_ssa * 8));
        dv.csa         = static_cast<CsaSuperDev const *>(up(d->csa_bv, d->n_csa_sb * 48));
        dv.seqs        = static_cast<unsigned char const *>(up(d->seqs, d->n_residues));
        dv.seqDelims   = static_cast<unsigned long long const *>(up(d->seq_delims, (d->n_seqs + 1) * 8));
        dv.origDelims  = dv.seqDelims;
        dv.sbjShift    = d->red_alph == LGPU_ALPH_DNA3BS ? 1 : 0;
        dv.bsMode      = d->red_alph == LGPU_ALPH_DNA3BS ? 1 : 0;
        dv.sbjFrames   = d->red_alph == LGPU_ALPH_DNA3BS ? 2 : 1;
        dv.nSeqs       = d->n_seqs;
        traceTime("index: device memory allocated");
        runUploads(pending, device);
        pending.clear();
        traceTime("index: main blobs uploaded");
        LGPU_CUDA(cudaMemcpyToSymbol(cDna5Translate, kDna5Translate, 125));
        // dbTotalLength = sum of reduced subject lengths (src/search_algo.hpp:317-318)
        ix->dbTotalLength = d->n_residues * (d->red_alph == LGPU_ALPH_DNA3BS ? 2 : 1);
        ix->sbjDelimsHost.assign(d->seq_delims, d->seq_delims + d->n_seqs + 1);
        if (d->trans_alph == LGPU_ALPH_AMINO_ACID && d->orig_alph == LGPU_ALPH_DNA5)
        {
            // translated subjects (TBLASTN / TBLASTX): transSbjSeqs = seqs | translate_join
            // (src/shared_definitions.hpp:246-255); made once here, 2 bytes per stored nucleotide
            std::vector<uint64_t> & td = ix->sbjDelimsHost;
            td.assign(d->n_seqs * 6 + 1, 0);
            for (uint64_t sq = 0; sq < d->n_seqs; ++sq)
                for (uint32_t f = 0; f < 6; ++f)
                    td[sq * 6 + f + 1] = td[sq * 6 + f] + translatedFrameLength(d->seq_delims[sq + 1] - d->seq_delims[sq], f);
            uint64_t const tTotal = td.back();
            unsigned long long * dTd = static_cast<unsigned long long *>(up(td.data(), td.size() * 8));
            unsigned char *      dT  = nullptr;
            LGPU_CUDA(cudaMalloc(&dT, std::max<uint64_t>(tTotal, 1)));
            ix->allocs.push_back(dT);
            ix->allocBytes.push_back(static_cast<size_t>(std::max<uint64_t>(tTotal, 1)));
            ix->bytes += tTotal;
            unsigned char * dComp = static_cast<unsigned char *>(up(kDna5Complement, 5));
            uint64_t const  nFr   = d->n_seqs * 6;
            runUploads(pending, device);
            pending.clear();
            translateSubjectsKernel<<<static_cast<unsigned int>(std::min<uint64_t>(std::max<uint64_t>(nFr, 1), 1u << 20)), 128>>>(
              dv.seqs, dv.origDelims, d->n_seqs, dComp, dTd, dT);
            LGPU_CUDA(cudaGetLastError());
            dv.seqs           = dT;
            dv.seqDelims      = dTd;
            dv.sbjFrames      = 6;
            ix->dbTotalLength = tTotal;
        }
        dv.nRows       = d->C[d->sigma];
        dv.bitsForPos  = static_cast<unsigned int>(d->bits_for_position);
        dv.posMask     = (1ull << d->bits_for_position) - 1;
        dv.blockBytes  = d->block_bytes;
        dv.planesOff   = d->planes_offset;
        dv.sigma       = d->sigma;
        dv.sigmaBits   = d->sigma_bits;
        dv.singleSuper = d->n_super == 1;
        for (unsigned int s = 0; s < 32; ++s)
            dv.Cbase[s] = 0;
        for (unsigned int s = 0; s < d->sigma; ++s)
            dv.Cbase[s] = d->C[s] + (dv.singleSuper ? d->super_blocks[s] : 0);
        ix->meta              = *d;
        ix->meta.occ_blocks   = nullptr;
        ix->meta.super_blocks = nullptr;
        ix->meta.C            = nullptr;
        ix->meta.ssa          = nullptr;
        ix->meta.csa_bv       = nullptr;
        ix->meta.seqs         = nullptr;
        ix->meta.seq_delims   = nullptr;
        ix->meta.ids          = nullptr;
        ix->meta.id_delims    = nullptr;
        traceTime("index: allocations done, uploading");
        runUploads(pending, device);
        LGPU_CUDA(cudaDeviceSynchronize());
        traceTime("index: uploaded");
        // re-pack the occurrence table: one aligned sector (or line) per rank; LAMBDA_B200_OCC_PACK=0 keeps the file's layout
        dv.occP = nullptr;
        dv.mid  = nullptr;
        char const * const pe = std::getenv("LAMBDA_B200_OCC_PACK");
        if ((!pe || std::atoi(pe) != 0) && d->n_blocks)
        {
            unsigned int const need = 8 * d->sigma_bits + 2 * (d->sigma - 1);
            dv.pCntOff              = 8 * d->sigma_bits;
            dv.pStride              = need <= 32 ? 32u : need <= 64 ? 64u : (need + 63) / 64 * 64;
            uint64_t const nGroups  = (d->n_blocks + 1023) / 1024;
            unsigned char *      dP = nullptr;
            unsigned long long * dM = nullptr;
            LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&dP), d->n_blocks * dv.pStride));
            ix->allocs.push_back(dP);
            ix->allocBytes.push_back(d->n_blocks * dv.pStride);
            LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&dM), nGroups * d->sigma * 8));
            ix->allocs.push_back(dM);
            ix->allocBytes.push_back(nGroups * d->sigma * 8);
            ix->bytes += d->n_blocks * dv.pStride + nGroups * d->sigma * 8;
            PackOccParams PP{};
            PP.ix      = dv;
            PP.nBlocks = d->n_blocks;
            for (unsigned int s2 = 0; s2 < d->sigma; ++s2)
                PP.C[s2] = d->C[s2];
            PP.occP = dP;
            PP.mid  = dM;
            packOccKernel<<<gridFor(d->n_blocks, 256), 256>>>(PP);
            LGPU_CUDA(cudaGetLastError());
            LGPU_CUDA(cudaDeviceSynchronize());
            dv.occP = dP;
            dv.mid  = dM;
        }
        traceTime("index: packed");
        char const * const ve = std::getenv("LAMBDA_B200_VALIDATE");
        if (!ve || std::atoi(ve) != 0)
        {
            // O(size) consistency checks of the uploaded blobs (kernels_fm.cuh validateIndexKernel)
            unsigned int * dFlags = nullptr;
            unsigned int   flags  = 0;
            LGPU_CUDA(cudaMalloc(reinterpret_cast<void **>(&dFlags), 4));
            LGPU_CUDA(cudaMemset(dFlags, 0, 4));
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            validateIndexKernel<<<static_cast<unsigned int>(sms) * 16, 256>>>(dv, d->n_blocks, d->n_ssa, d->n_csa_sb,
                                                                                d->n_seqs * dv.sbjFrames, dFlags);
            cudaError_t const e1 = cudaMemcpy(&flags, dFlags, 4, cudaMemcpyDeviceToHost);
            cudaFree(dFlags);
            LGPU_CUDA(e1);
            if (flags)
                throw LbaError(std::string("index file corrupt:") + ((flags & 1u) ? " occ block counts are inconsistent;" : "") +
                               ((flags & 2u) ? " CSA bit vector ranks outside the sampled suffix array;" : "") +
                               ((flags & 4u) ? " sampled suffix array names a sequence / position that does not exist;" : ""));
        }
        *out = ix.release();
    });
}

int lgpu_device_warmup(int device)
{
    return guarded(nullptr, [&] {
        int nDev = 0;
        if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
            throw CudaError("no CUDA device available (lambda_b200 has no CPU fallback)");
        LGPU_CUDA(cudaSetDevice(device));
        LGPU_CUDA(cudaFree(nullptr)); // creates the context
    });
}

int lgpu_index_clone(lgpu_index ** out, lgpu_index const * src, int device)
{
    if (!out || !src)
        return LGPU_ERR_ARG;
    *out = nullptr;
    return guarded(nullptr, [&] {
        LGPU_CUDA(cudaSetDevice(device));
        auto ix           = std::make_unique<lgpu_index>();
        ix->device        = device;
        ix->meta          = src->meta;
        ix->bytes         = src->bytes;
        ix->dbTotalLength = src->dbTotalLength;
        ix->sbjDelimsHost = src->sbjDelimsHost;
        ix->dev           = src->dev;
        if (device != src->device)
        {
            int can = 0;
            LGPU_CUDA(cudaDeviceCanAccessPeer(&can, device, src->device));
            if (can)
            {
                cudaError_t const e = cudaDeviceEnablePeerAccess(src->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    LGPU_CUDA(e);
                cudaGetLastError();
            }
        }
        // every blob of the source, copied GPU to GPU (NVLink when the devices are peers)
        for (size_t k = 0; k < src->allocs.size(); ++k)
        {
            void * dp = nullptr;
            LGPU_CUDA(cudaMalloc(&dp, std::max<size_t>(src->allocBytes[k], 1)));
            ix->allocs.push_back(dp);
            ix->allocBytes.push_back(src->allocBytes[k]);
            if (src->allocBytes[k])
                LGPU_CUDA(cudaMemcpyPeerAsync(dp, device, src->allocs[k], src->device, src->allocBytes[k], nullptr));
        }
        auto remap = [&](void const * p) -> void const * {
            if (!p)
                return nullptr;
            for (size_t k = 0; k < src->allocs.size(); ++k)
            {
                char const * b = static_cast<char const *>(src->allocs[k]);
                char const * q = static_cast<char const *>(p);
                if (q >= b && q < b + std::max<size_t>(src->allocBytes[k], 1))
                    return static_cast<char const *>(ix->allocs[k]) + (q - b);
            }
            throw CudaError("lgpu_index_clone: a device pointer of the source index is not inside its allocations");
        };
        DevIndex & dv = ix->dev;
        dv.occ        = static_cast<unsigned char const *>(remap(dv.occ));
        dv.super      = static_cast<unsigned long long const *>(remap(dv.super));
        dv.ssa        = static_cast<unsigned long long const *>(remap(dv.ssa));
        dv.csa        = static_cast<CsaSuperDev const *>(remap(dv.csa));
        dv.seqs       = static_cast<unsigned char const *>(remap(dv.seqs));
        dv.seqDelims  = static_cast<unsigned long long const *>(remap(dv.seqDelims));
        dv.origDelims = static_cast<unsigned long long const *>(remap(dv.origDelims));
        dv.occP       = static_cast<unsigned char const *>(remap(dv.occP));
        dv.mid        = static_cast<unsigned long long const *>(remap(dv.mid));
        LGPU_CUDA(cudaMemcpyToSymbol(cDna5Translate, kDna5Translate, 125));
        LGPU_CUDA(cudaDeviceSynchronize());
        traceTime("index: validated");
        *out = ix.release();
    });
}

void lgpu_index_destroy(lgpu_index * ix) { delete ix; }

uint64_t lgpu_index_device_bytes(lgpu_index const * ix) { return ix ? ix->bytes : 0; }
uint64_t lgpu_index_db_total_length(lgpu_index const * ix) { return ix ? ix->dbTotalLength : 0; }
uint64_t lgpu_index_db_num_seqs(lgpu_index const * ix) { return ix ? ix->meta.n_seqs : 0; }

int lgpu_params_default(lgpu_params * out, uint32_t domain, char const * profile)
{
    if (!out)
        return LGPU_ERR_ARG;
    return paramsDefault(*out, domain, profile);
}

int lgpu_ctx_create(lgpu_ctx ** out, lgpu_index const * ix, lgpu_params const * p)
{
    if (!out || !ix || !p)
        return LGPU_ERR_ARG;
    *out = nullptr;
    return guarded(nullptr, [&] {
        auto c = makeContext(ix, *p);
        // Sub-batches in flight per search call (worker contexts are created on first use).  With the records finalised
        // on the device the host part of a protein step is under a millisecond and one stream is fastest (searchp
        // 51.7 ms against 53.8 ms with three); a short-read step still has ~10 ms of host work per million records to
        // hide (searchn 55.5 ms with four against 61.4 ms with one) -- profiles/r2_sweep_streams.jsonl.
        c->streams = p->domain == LGPU_DOMAIN_PROTEIN ? 1 : 4;
        if (char const * e = std::getenv("LAMBDA_B200_STREAMS"))
            c->streams = static_cast<unsigned int>(std::max(1, std::min(8, std::atoi(e))));
        if (char const * e = std::getenv("LAMBDA_B200_MIN_SUBBATCH"))
            c->minSubBatch = static_cast<uint64_t>(std::max(1, std::atoi(e)));
        *out = c.release();
    });
}

void lgpu_ctx_destroy(lgpu_ctx * c) { delete c; }

int lgpu_ctx_set_streams(lgpu_ctx * c, uint32_t n)
{
    if (!c || n < 1 || n > 8)
        return LGPU_ERR_ARG;
    c->streams = n;
    return LGPU_OK;
}

char const * lgpu_last_error(lgpu_ctx const * c) { return (c && !c->err.empty()) ? c->err.c_str() : g_lastError.c_str(); }

int lgpu_search_batch(lgpu_ctx * c, lgpu_query_batch const * q, lgpu_hits * out, lgpu_stats * stats)
{
    if (!c || !q || !out)
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] { searchBatch(*c, *q, out, stats); });
}

int lgpu_ctx_export_hits(lgpu_ctx * c, void * devDst, uint64_t capRecords, uint64_t firstQuery, uint64_t * nOut)
{
    if (!c || !nOut || (capRecords && !devDst))
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] {
        LGPU_CUDA(cudaSetDevice(c->index->device));
        uint64_t total = 0;
        for (auto const & part : c->lastParts)
            total += part.n;
        *nOut = total;
        if (total > capRecords)
            return; // the caller learns the size it needs and calls again
        lgpu_hit * dst = static_cast<lgpu_hit *>(devDst);
        uint64_t   off = 0;
        for (auto const & part : c->lastParts)
        {
            if (part.n)
                exportHitsKernel<<<gridFor(part.n, 256), 256, 0, c->stream>>>(part.ctx->finalDev, part.n,
                                                                               static_cast<unsigned int>(firstQuery + part.qBase),
                                                                               static_cast<unsigned int>(part.cigarBase), dst + off);
            off += part.n;
        }
        LGPU_CUDA(cudaGetLastError());
        syncStream(*c);
    });
}

int lgpu_hits_fill_scores(lgpu_ctx * c, lgpu_hit * hits, uint64_t n)
{
    if (!c || (n && !hits))
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] {
        if (!c->evc)
            c->evc = std::make_unique<EValueComputer>(c->scoring.ka, c->index->dbTotalLength, c->di.qIsTranslated);
        for (uint64_t i = 0; i < n; ++i)
        {
            hits[i].bit_score = c->evc->bitsCached(hits[i].score);
            hits[i].evalue    = c->evc->evalueCached(hits[i].score, hits[i].q_len);
        }
    });
}

int lgpu_seed_batch(lgpu_ctx * c, lgpu_query_batch const * q, int phase, lgpu_match const ** matches, uint64_t * n,
                    lgpu_stats * stats)
{
    if (!c || !q || !matches || !n || (phase != 1 && phase != 2))
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] {
        c->timersUsed = 0;
        LGPU_CUDA(cudaSetDevice(c->index->device));
        {
            std::vector<uint64_t> offsTmp;
            uploadQueries(*c, viewOf(*c, *q, offsTmp), stats);
        }
        unsigned int const nQ = static_cast<unsigned int>(c->nQueries);
        c->dActive.reserve(nQ);
        if (nQ)
            iotaKernel<<<gridFor(nQ, 256), 256, 0, c->stream>>>(c->dActive.p, nQ);
        uint64_t const maxLen = c->maxQueryLen;
        uint64_t const nM = runSeeding(*c, phase == 1 ? c->params.opts0 : c->params.opts, c->dActive.p,
                                       nQ, static_cast<unsigned int>(maxLen), stats);
        c->matchesHost.resize(nM);
        if (nM)
            LGPU_CUDA(cudaMemcpyAsync(c->matchesHost.data(), c->dMatches.p, nM * sizeof(lgpu_match), cudaMemcpyDeviceToHost,
                                      c->stream));
        syncStream(*c);
        resolveTimers(*c);
        *matches = c->matchesHost.data();
        *n       = nM;
    });
}

int lgpu_merge_matches(lgpu_ctx * c, lgpu_query_batch const * q, lgpu_match const * in, uint64_t nIn,
                       lgpu_match const ** merged, uint64_t * nMerged, lgpu_stats * stats)
{
    if (!c || !q || !merged || !nMerged || (nIn && !in))
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] {
        c->timersUsed = 0;
        LGPU_CUDA(cudaSetDevice(c->index->device));
        {
            std::vector<uint64_t> offsTmp;
            uploadQueries(*c, viewOf(*c, *q, offsTmp), stats);
        }
        c->dUserMatches.reserve(nIn);
        if (nIn)
            LGPU_CUDA(cudaMemcpyAsync(c->dUserMatches.p, in, nIn * sizeof(lgpu_match), cudaMemcpyHostToDevice, c->stream));
        uint64_t const nOut = runMerge(*c, c->dUserMatches.p, nIn, stats);
        c->matchesHost.resize(nOut);
        if (nOut)
            LGPU_CUDA(cudaMemcpyAsync(c->matchesHost.data(), c->dMerged.p, nOut * sizeof(lgpu_match), cudaMemcpyDeviceToHost,
                                      c->stream));
        syncStream(*c);
        resolveTimers(*c);
        *merged  = c->matchesHost.data();
        *nMerged = nOut;
    });
}

static void checkWindows(lgpu_ctx & c, lgpu_match const * w, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i)
    {
        uint64_t const q = w[i].qry_id / c.di.qryNumFrames;
        if (q >= c.nQueries || w[i].subj_id / c.di.sbjNumFrames >= c.index->meta.n_seqs)
            throw ArgError("window refers to a query/subject that does not exist");
        uint64_t       qLen = c.qOffsHost[q + 1] - c.qOffsHost[q];
        if (c.di.qIsTranslated)
            qLen = translatedFrameLength(qLen, w[i].qry_id % 6);
        uint64_t const sIdx = w[i].subj_id >> c.index->dev.sbjShift;
        uint64_t const sLen = c.index->sbjDelimsHost[sIdx + 1] - c.index->sbjDelimsHost[sIdx];
        if (w[i].qry_start > w[i].qry_end || w[i].qry_end > qLen || w[i].subj_start > w[i].subj_end ||
            w[i].subj_end > sLen)
            throw ArgError("window coordinates out of range");
    }
}

int lgpu_extend_scores(lgpu_ctx * c, lgpu_query_batch const * q, lgpu_match const * win, uint64_t n, int32_t * scores,
                       lgpu_stats * stats)
{
    if (!c || !q || (n && (!win || !scores)))
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] {
        c->timersUsed = 0;
        LGPU_CUDA(cudaSetDevice(c->index->device));
        {
            std::vector<uint64_t> offsTmp;
            uploadQueries(*c, viewOf(*c, *q, offsTmp), stats);
        }
        if (n >= (1ull << 31))
            throw ArgError("more than 2^31 windows in one call");
        if (n == 0)
            return;
        checkWindows(*c, win, n);
        c->dUserMatches.reserve(n);
        LGPU_CUDA(cudaMemcpyAsync(c->dUserMatches.p, win, n * sizeof(lgpu_match), cudaMemcpyHostToDevice, c->stream));
        c->dScores.reserve(n);
        runScorePass(*c, c->dUserMatches.p, static_cast<unsigned int>(n), c->dScores.p, stats);
        LGPU_CUDA(cudaMemcpyAsync(scores, c->dScores.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
        syncStream(*c);
        resolveTimers(*c);
        checkScoreOverflow(*c);
    });
}

int lgpu_extend_trace(lgpu_ctx * c, lgpu_query_batch const * q, lgpu_match const * win, uint64_t n, lgpu_hit * out,
                      lgpu_stats * stats)
{
    if (!c || !q || (n && (!win || !out)))
        return LGPU_ERR_ARG;
    return guarded(&c->err, [&] {
        c->timersUsed = 0;
        LGPU_CUDA(cudaSetDevice(c->index->device));
        {
            std::vector<uint64_t> offsTmp;
            uploadQueries(*c, viewOf(*c, *q, offsTmp), stats);
        }
        if (n >= (1ull << 31))
            throw ArgError("more than 2^31 windows in one call");
        if (n == 0)
            return;
        checkWindows(*c, win, n);
        c->dUserMatches.reserve(n);
        LGPU_CUDA(cudaMemcpyAsync(c->dUserMatches.p, win, n * sizeof(lgpu_match), cudaMemcpyHostToDevice, c->stream));
        c->cigar.clear(); // with want_cigar the records index into the context's run buffer of THIS call
        runTracePass(*c, c->dUserMatches.p, static_cast<unsigned int>(n), stats);
        LGPU_CUDA(cudaMemcpyAsync(out, c->dHits.p, n * sizeof(lgpu_hit), cudaMemcpyDeviceToHost, c->stream));
        syncStream(*c);
        resolveTimers(*c);
        checkScoreOverflow(*c);
    });
}

int lgpu_fm_rank(lgpu_index const * ix, uint64_t const * idx, uint8_t const * symb, uint64_t n, uint64_t * out)
{
    if (!ix || (n && (!idx || !symb || !out)))
        return LGPU_ERR_ARG;
    return guarded(nullptr, [&] {
        LGPU_CUDA(cudaSetDevice(ix->device));
        for (uint64_t i = 0; i < n; ++i)
            if (idx[i] > ix->dev.nRows || symb[i] >= ix->dev.sigma)
                throw ArgError("rank query out of range");
        DevBuf<unsigned long long> dIdx, dOut;
        DevBuf<unsigned char>      dSym;
        dIdx.reserve(n);
        dOut.reserve(n);
        dSym.reserve(n);
        LGPU_CUDA(cudaMemcpy(dIdx.p, idx, n * 8, cudaMemcpyHostToDevice));
        LGPU_CUDA(cudaMemcpy(dSym.p, symb, n, cudaMemcpyHostToDevice));
        fmRankKernel<<<gridFor(n, 128), 128>>>(ix->dev, dIdx.p, dSym.p, n, dOut.p);
        LGPU_CUDA(cudaGetLastError());
        LGPU_CUDA(cudaMemcpy(out, dOut.p, n * 8, cudaMemcpyDeviceToHost));
    });
}

int lgpu_fm_locate(lgpu_index const * ix, uint64_t const * rows, uint64_t n, uint64_t * subj, uint64_t * pos)
{
    if (!ix || (n && (!rows || !subj || !pos)))
        return LGPU_ERR_ARG;
    return guarded(nullptr, [&] {
        LGPU_CUDA(cudaSetDevice(ix->device));
        for (uint64_t i = 0; i < n; ++i)
            if (rows[i] >= ix->dev.nRows)
                throw ArgError("locate row out of range");
        DevBuf<unsigned long long> dRows, dSubj, dPos;
        dRows.reserve(n);
        dSubj.reserve(n);
        dPos.reserve(n);
        LGPU_CUDA(cudaMemcpy(dRows.p, rows, n * 8, cudaMemcpyHostToDevice));
        fmLocateKernel<<<gridFor(n, 128), 128>>>(ix->dev, dRows.p, n, dSubj.p, dPos.p);
        LGPU_CUDA(cudaGetLastError());
        LGPU_CUDA(cudaMemcpy(subj, dSubj.p, n * 8, cudaMemcpyDeviceToHost));
        LGPU_CUDA(cudaMemcpy(pos, dPos.p, n * 8, cudaMemcpyDeviceToHost));
    });
}

int lgpu_bit_score(lgpu_params const * p, int32_t raw, double * out)
{
    if (!p || !out)
        return LGPU_ERR_ARG;
    KarlinAltschul const ka = selectKA(*p);
    if (!ka.valid)
        return LGPU_ERR_ARG;
    *out = bitScore(ka, raw);
    return LGPU_OK;
}

int lgpu_ka_params(lgpu_params const * p, double * lambda, double * k, double * h)
{
    if (!p || !lambda || !k || !h)
        return LGPU_ERR_ARG;
    KarlinAltschul const ka = selectKA(*p);
    if (!ka.valid)
        return LGPU_ERR_ARG;
    *lambda = ka.lambda;
    *k      = ka.K;
    *h      = ka.H;
    return LGPU_OK;
}

int lgpu_score_matrix(lgpu_params const * p, int8_t * out)
{
    if (!p || !out)
        return LGPU_ERR_ARG;
    Scoring sc;
    if (makeScoring(sc, *p) != 0)
        return LGPU_ERR_ARG;
    std::memcpy(out, sc.matrix, 32 * 32);
    return LGPU_OK;
}

int lgpu_evalue(lgpu_params const * p, int32_t raw, uint64_t qLen, uint64_t dbLen, double * out)
{
    if (!p || !out)
        return LGPU_ERR_ARG;
    KarlinAltschul const ka = selectKA(*p);
    if (!ka.valid)
        return LGPU_ERR_ARG;
    EValueComputer ev(ka, dbLen, domainInfo(p->domain, 0, p->query_alph).qIsTranslated);
    *out = ev.evalue(raw, qLen);
    return LGPU_OK;
}

int lgpu_min_raw_score(lgpu_params const * p, uint64_t qLen, uint64_t dbLen, int32_t * out)
{
    if (!p || !out)
        return LGPU_ERR_ARG;
    KarlinAltschul const ka = selectKA(*p);
    if (!ka.valid)
        return LGPU_ERR_ARG;
    EValueComputer        ev(ka, dbLen, domainInfo(p->domain, 0, p->query_alph).qIsTranslated);
    ScoreThresholds const t = scoreThresholds(*p, ev, qLen);
    *out                    = std::max(t.minBit, t.minEval);
    return LGPU_OK;
}

int lgpu_format_m8(lgpu_params const * p, lgpu_hit const * h, char const * qId, char const * sId, char * buf, size_t cap)
{
    if (!p || !h || !qId || !sId || !buf)
        return LGPU_ERR_ARG;
    return formatM8(p->domain, *h, qId, std::strlen(qId), sId, std::strlen(sId), buf, cap);
}

int lgpu_tabular_column(char const * label)
{
    if (!label)
        return -1;
    for (int i = 0; i < kNumTabColumns; ++i)
        if (!std::strcmp(label, tabColumnOptionLabels()[i]))
            return i;
    return -1;
}

char const * lgpu_tabular_column_name(uint32_t column)
{
    return column < static_cast<uint32_t>(kNumTabColumns) ? tabColumnOptionLabels()[column] : nullptr;
}

char const * lgpu_tabular_column_label(uint32_t column)
{
    return column < static_cast<uint32_t>(kNumTabColumns) ? tabColumnLabels()[column] : nullptr;
}

int lgpu_tabular_column_supported(uint32_t column)
{
    return column < static_cast<uint32_t>(kNumTabColumns) && column != TAB_S_TAX_IDS && column != TAB_LCA_ID &&
           column != TAB_LCA_TAX_ID;
}

int lgpu_tabular_column_implemented(uint32_t column) { return tabColumnImplemented(column) ? 1 : 0; }

int lgpu_format_tabular(lgpu_params const * p, lgpu_hit const * h, char const * qId, char const * sId, uint32_t const * columns,
                        size_t nColumns, char * buf, size_t cap)
{
    if (!p || !h || !qId || !sId || !buf || (!columns && nColumns))
        return LGPU_ERR_ARG;
    int const n = formatTabular(p->domain, *h, qId, std::strlen(qId), sId, std::strlen(sId), columns, nColumns, buf, cap);
    return n < 0 ? LGPU_ERR_ARG : n;
}

} // extern "C"
