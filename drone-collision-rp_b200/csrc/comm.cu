// comm.cu -- multi-process exchange of the per-pair hit counts (the loop it shards is
// /root/reference/Monte_Carlo.cpp:28-34; the counter it combines is Drone.cpp:276).
//
// One process per GPU.  dcrp_engine_attach_comm builds an NCCL communicator from a 128-byte unique id the
// caller ships to every rank (NCCL is dlopen'ed: libdcrp.so has no link-time dependency on it, and inside a
// torch process it binds to the copy torch already loaded).  NCCL does the bootstrap -- an all-gather of
// CUDA IPC handles -- and is the fallback collective; the per-step all-reduce itself is two tiny kernels of
// our own over NVLink peer memory:
//   k_comm_push  one CTA per peer: stores this rank's counts into its slot of the peer's mailbox
//                (remote stores through the IPC mapping), system-scope fence, then the slot's sequence flag;
//   k_comm_sum   one CTA: waits (bounded) for every rank's flag of this sequence number, sums the slots.
// Why not ncclAllReduce per step: the trial kernels are persistent grids that fill every SM's register file,
// so NCCL's 512-thread kernel cannot become resident beside them and its "overlapped" all-reduce ran exposed
// in some runs (round 1: +50 us of a 300 us step at N=4).  These kernels need 2 K registers and always fit.
// Two mailbox parities make the exchange safe without a barrier: rank A can only push sequence s+2 after it
// has summed s+1, which needed B's push of s+1, which B issued after its own sum of s (stream order).
#include <dlfcn.h>
#include <nccl.h>

namespace dcrp {

constexpr int COMM_MAX_RANKS = 64;
constexpr size_t COMM_MAX_PAIRS = 1u << 16;        // counters per rank slot (512 KB); larger scenarios use NCCL

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;

    bool load()
    {
        if (handle) return true;
        // the copy already in the process (torch bundles its own libnccl.so.2) wins; else the system one
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(handle, name);
            if (!p) error = std::string("libnccl lacks ") + name;
            return p;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRankConfig = reinterpret_cast<decltype(CommInitRankConfig)>(sym("ncclCommInitRankConfig"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
        if (!GetUniqueId || !CommInitRankConfig || !CommDestroy || !AllReduce || !AllGather || !GetErrorString) {
            handle = nullptr;
            return false;
        }
        return true;
    }
};

static NcclApi& nccl_api()
{
    static NcclApi api;
    return api;
}

// Mailbox (device memory of the owning rank, mapped by every peer):
//   flags[2][COMM_MAX_RANKS] sequence numbers, then slots[2][nranks][COMM_MAX_PAIRS] counters
struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    bool p2p = false;                       // peer mailboxes mapped: the exchange runs on our own kernels
    unsigned long long* mailbox = nullptr;  // local
    unsigned long long* peer[COMM_MAX_RANKS] = {};     // peer[r] = rank r's mailbox as mapped here (peer[rank] = mailbox)
    unsigned long long seq = 0;
    uint32_t* h_err = nullptr;              // mapped pinned word, set by k_comm_sum when a peer did not arrive in time
    uint32_t* d_err = nullptr;              // its device alias
    int nccl_version = 0;
    uint64_t n_p2p = 0, n_nccl = 0;         // all-reduces done by each route
};

__host__ __device__ inline size_t comm_flag_words() { return 2 * (size_t)COMM_MAX_RANKS; }
__host__ __device__ inline size_t comm_slot_off(int parity, int nranks, int r)
{
    return comm_flag_words() + ((size_t)parity * (size_t)nranks + (size_t)r) * COMM_MAX_PAIRS;
}
inline size_t comm_mailbox_bytes(int nranks) { return sizeof(unsigned long long) * comm_slot_off(2, nranks, 0); }

struct CommPeers { unsigned long long* p[COMM_MAX_RANKS]; };

// blockIdx.x = k-th peer (skipping this rank).  Remote stores, then the flag, release at system scope.
__global__ void __launch_bounds__(256) k_comm_push(const CommPeers peers, int rank, int nranks, int parity,
                                                   unsigned long long seq, const unsigned long long* __restrict__ counts,
                                                   uint32_t n)
{
    int dst = (int)blockIdx.x;
    if (dst >= rank) ++dst;
    DCRP_ASSERT(dst < nranks && n <= COMM_MAX_PAIRS && peers.p[dst] != nullptr);
    unsigned long long* mb = peers.p[dst];
    unsigned long long* slot = mb + comm_slot_off(parity, nranks, rank);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) slot[i] = counts[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(mb + (size_t)parity * COMM_MAX_RANKS + rank), "l"(seq)
                     : "memory");
    }
}

// counts[i] = sum over ranks (own counts + the peers' slots), in place.  Waits at most timeout_ns per peer.
__global__ void __launch_bounds__(256) k_comm_sum(unsigned long long* mailbox, int rank, int nranks, int parity,
                                                  unsigned long long seq, unsigned long long* counts, uint32_t n,
                                                  unsigned long long timeout_ns, uint32_t* err)
{
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < nranks && (int)threadIdx.x != rank) {
        const unsigned long long* flag = mailbox + (size_t)parity * COMM_MAX_RANKS + threadIdx.x;
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= seq) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { ok = 0; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) { *reinterpret_cast<volatile uint32_t*>(err) = 1u; __threadfence_system(); }
        return;
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned long long s = counts[i];
        for (int r = 0; r < nranks; ++r)
            if (r != rank) s += __ldcv(&mailbox[comm_slot_off(parity, nranks, r) + i]);
        counts[i] = s;
    }
}

}  // namespace dcrp
