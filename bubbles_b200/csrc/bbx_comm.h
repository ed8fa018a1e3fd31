// Slab-neighbour transports of the bbx engine (host side).
//
// The reference has no multi-GPU code at all (no cudaSetDevice, no peer access, no
// NCCL/MPI); this layer is new.  A slab engine talks to rank-1 ("lo", smaller z) and rank+1 ("hi"):
//   exchange()          contiguous device ranges to / from the two neighbours (ghost planes, migrants)
//   allreduce_max_u32() global flags / maxima (big-move rule, CFL force maximum, density error)
//   neighbor_counts()   host integers to / from each neighbour (sizes of the next exchange, owned counts)
//   share_arrays()      once, after the communicator is up: the neighbours' device pointers to the arrays
//                       that carry ghost slots (CUDA IPC between processes, plain pointers in-process), so
//                       that the sweeps can store boundary-plane results straight into the neighbour's ghost
//                       slots over NVLink instead of a send / recv pair per phase
// Two implementations:
//   NcclComm   one process per GPU, ncclSend/ncclRecv groups over NVLink (libnccl is dlopen'ed so that
//              the single-GPU library has no link-time dependency on it)
//   LocalComm  several slab engines of ONE process on ONE device, one host thread per engine, peer
//              copies with cudaMemcpyAsync + events.  Same engine code path; lets the slab logic be
//              checked bit-exactly against the single-domain engine on a single GPU.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <chrono>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>

struct BbxSeg { void *ptr; size_t bytes; };
#define BBX_MAX_SEGS 8
#define BBX_LOCAL_MAX_RANKS 64
#define BBX_PEER_NPTR 11  // pos[2], vel[2], rec, pred, posq, halo flags + mailbox, pid[2], ghost cell-table slices
#define BBX_NCOUNTS 2     // integers per neighbour in neighbor_counts: boundary-plane particles, owned particles
// allocation bases of the arrays a neighbour may write into, and the ghost slots in front of slot 0
struct BbxPeerArrays { void *p[BBX_PEER_NPTR]; long long gc; };

struct BbxComm {
    int rank = 0, nranks = 1;
    std::string err;
    virtual ~BbxComm(){}
    bool has_lo() const { return rank > 0; }
    bool has_hi() const { return rank + 1 < nranks; }
    // send_lo[k] of this rank pairs with recv_hi[k] of rank-1; send_hi[k] with recv_lo[k] of rank+1
    virtual int exchange(cudaStream_t s, const BbxSeg *send_lo, const BbxSeg *recv_lo, int n_lo,
                         const BbxSeg *send_hi, const BbxSeg *recv_hi, int n_hi) = 0;
    virtual int allreduce_max_u32(cudaStream_t s, unsigned *dev, int count) = 0;
    // returns with the host values filled in (synchronises the stream); BBX_NCOUNTS integers each
    virtual int neighbor_counts(cudaStream_t s, const int *to_lo, const int *to_hi, int *from_lo, int *from_hi) = 0;
    // *ok = 1: lo / hi hold pointers valid on THIS device / in THIS process for the neighbours' arrays
    // (a missing neighbour's entry is zeroed); *ok = 0 on every rank if any rank cannot map its neighbours
    virtual int share_arrays(cudaStream_t s, const BbxPeerArrays &mine, int want, BbxPeerArrays *lo, BbxPeerArrays *hi, int *ok) = 0;
    virtual int barrier() = 0;
    // several engines of ONE process on ONE device (test transport): their streams share the device's few hardware queues,
    // so a backlog of spinning halo waits from one engine can sit in front of the very kernel it waits for -- such engines
    // drain their stream at the end of every sub-step (one process per GPU never needs to)
    virtual bool shares_device() const { return false; }
};

// ------------------------------------------------------------------------------------------ NCCL
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string load_error;
    bool load(){
        if(handle) return true;
        // a libnccl already mapped by the host process (e.g. the one bundled with PyTorch) is reused:
        // the loader matches the soname
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for(const char *n : names){ handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if(handle) break; }
        if(!handle){ load_error = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define BBX_NCCL_SYM(field, name) do{ *(void **)(&field) = dlsym(handle, name); if(!field){ load_error = std::string("libnccl lacks ") + name; handle = nullptr; return false; } }while(0)
        BBX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        BBX_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        BBX_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        BBX_NCCL_SYM(Send, "ncclSend");
        BBX_NCCL_SYM(Recv, "ncclRecv");
        BBX_NCCL_SYM(AllReduce, "ncclAllReduce");
        BBX_NCCL_SYM(GroupStart, "ncclGroupStart");
        BBX_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        BBX_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef BBX_NCCL_SYM
        return true;
    }
};
static NcclApi g_nccl;

#define BBX_NCCL(call) do{ ncclResult_t _r = (call); if(_r != ncclSuccess){ err = std::string(#call) + ": " + g_nccl.GetErrorString(_r); return 1; } }while(0)
#define BBX_CUC(call) do{ cudaError_t _e = (call); if(_e != cudaSuccess){ err = std::string(#call) + ": " + cudaGetErrorString(_e); return 1; } }while(0)

struct NcclComm : BbxComm {
    ncclComm_t comm = nullptr;
    int *xdev = nullptr;   // [0..K) to lo, [K..2K) to hi, [2K..3K) from lo, [3K..4K) from hi  (K = BBX_NCOUNTS)
    int *xhost = nullptr;  // pinned
    void *opened[2 * BBX_PEER_NPTR]; int n_opened = 0; // IPC mappings of the neighbours' arrays
    int init(int rank_, int nranks_, const unsigned char *id){
        rank = rank_; nranks = nranks_;
        if(!g_nccl.load()){ err = g_nccl.load_error; return 1; }
        ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
        BBX_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
        BBX_CUC(cudaMalloc((void **)&xdev, 4 * BBX_NCOUNTS * sizeof(int)));
        BBX_CUC(cudaMallocHost((void **)&xhost, 4 * BBX_NCOUNTS * sizeof(int)));
        return 0;
    }
    ~NcclComm() override {
        for(int k = 0; k < n_opened; k++) cudaIpcCloseMemHandle(opened[k]);
        if(comm) g_nccl.CommDestroy(comm);
        if(xdev) cudaFree(xdev);
        if(xhost) cudaFreeHost(xhost);
    }
    int exchange(cudaStream_t s, const BbxSeg *send_lo, const BbxSeg *recv_lo, int n_lo,
                 const BbxSeg *send_hi, const BbxSeg *recv_hi, int n_hi) override {
        BBX_NCCL(g_nccl.GroupStart());
        if(has_lo()) for(int k = 0; k < n_lo; k++){
            if(send_lo[k].bytes) BBX_NCCL(g_nccl.Send(send_lo[k].ptr, send_lo[k].bytes, ncclChar, rank - 1, comm, s));
            if(recv_lo[k].bytes) BBX_NCCL(g_nccl.Recv(recv_lo[k].ptr, recv_lo[k].bytes, ncclChar, rank - 1, comm, s));
        }
        if(has_hi()) for(int k = 0; k < n_hi; k++){
            if(send_hi[k].bytes) BBX_NCCL(g_nccl.Send(send_hi[k].ptr, send_hi[k].bytes, ncclChar, rank + 1, comm, s));
            if(recv_hi[k].bytes) BBX_NCCL(g_nccl.Recv(recv_hi[k].ptr, recv_hi[k].bytes, ncclChar, rank + 1, comm, s));
        }
        BBX_NCCL(g_nccl.GroupEnd());
        return 0;
    }
    int allreduce_max_u32(cudaStream_t s, unsigned *dev, int count) override {
        BBX_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclUint32, ncclMax, comm, s));
        return 0;
    }
    int neighbor_counts(cudaStream_t s, const int *to_lo, const int *to_hi, int *from_lo, int *from_hi) override {
        const int K = BBX_NCOUNTS;
        for(int k = 0; k < K; k++){ xhost[k] = to_lo[k]; xhost[K + k] = to_hi[k]; xhost[2 * K + k] = 0; xhost[3 * K + k] = 0; }
        BBX_CUC(cudaMemcpyAsync(xdev, xhost, 4 * K * sizeof(int), cudaMemcpyHostToDevice, s));
        BBX_NCCL(g_nccl.GroupStart());
        if(has_lo()){
            BBX_NCCL(g_nccl.Send(xdev + 0, K * sizeof(int), ncclChar, rank - 1, comm, s));
            BBX_NCCL(g_nccl.Recv(xdev + 2 * K, K * sizeof(int), ncclChar, rank - 1, comm, s));
        }
        if(has_hi()){
            BBX_NCCL(g_nccl.Send(xdev + K, K * sizeof(int), ncclChar, rank + 1, comm, s));
            BBX_NCCL(g_nccl.Recv(xdev + 3 * K, K * sizeof(int), ncclChar, rank + 1, comm, s));
        }
        BBX_NCCL(g_nccl.GroupEnd());
        BBX_CUC(cudaMemcpyAsync(xhost + 2 * K, xdev + 2 * K, 2 * K * sizeof(int), cudaMemcpyDeviceToHost, s));
        BBX_CUC(cudaStreamSynchronize(s));
        for(int k = 0; k < K; k++){ from_lo[k] = has_lo() ? xhost[2 * K + k] : 0; from_hi[k] = has_hi() ? xhost[3 * K + k] : 0; }
        return 0;
    }
    // CUDA IPC: every rank exports its allocations, the handles travel to the two neighbours over NCCL, and
    // each rank maps its neighbours' allocations (NVLink peer access).  Agreement: a global max over a
    // "failed" flag, so either every rank stores into its neighbours' ghost slots or none does.
    int share_arrays(cudaStream_t s, const BbxPeerArrays &mine, int want, BbxPeerArrays *lo, BbxPeerArrays *hi, int *ok) override {
        struct Msg { cudaIpcMemHandle_t h[BBX_PEER_NPTR]; long long gc; int good; int pad; };
        memset(lo, 0, sizeof(*lo)); memset(hi, 0, sizeof(*hi));
        Msg *hm = nullptr, *dm = nullptr; // [0] mine, [1] from lo, [2] from hi
        BBX_CUC(cudaMallocHost((void **)&hm, 3 * sizeof(Msg)));
        BBX_CUC(cudaMalloc((void **)&dm, 3 * sizeof(Msg)));
        memset(hm, 0, 3 * sizeof(Msg));
        unsigned failed = want ? 0u : 1u;
        for(int k = 0; k < BBX_PEER_NPTR && !failed; k++) if(cudaIpcGetMemHandle(&hm[0].h[k], mine.p[k]) != cudaSuccess){ cudaGetLastError(); failed = 1u; }
        hm[0].gc = mine.gc; hm[0].good = failed ? 0 : 1;
        BBX_CUC(cudaMemcpyAsync(dm, hm, 3 * sizeof(Msg), cudaMemcpyHostToDevice, s));
        BBX_NCCL(g_nccl.GroupStart());
        if(has_lo()){
            BBX_NCCL(g_nccl.Send(dm + 0, sizeof(Msg), ncclChar, rank - 1, comm, s));
            BBX_NCCL(g_nccl.Recv(dm + 1, sizeof(Msg), ncclChar, rank - 1, comm, s));
        }
        if(has_hi()){
            BBX_NCCL(g_nccl.Send(dm + 0, sizeof(Msg), ncclChar, rank + 1, comm, s));
            BBX_NCCL(g_nccl.Recv(dm + 2, sizeof(Msg), ncclChar, rank + 1, comm, s));
        }
        BBX_NCCL(g_nccl.GroupEnd());
        BBX_CUC(cudaMemcpyAsync(hm, dm, 3 * sizeof(Msg), cudaMemcpyDeviceToHost, s));
        BBX_CUC(cudaStreamSynchronize(s));
        for(int side = 0; side < 2 && !failed; side++){
            if(side == 0 ? !has_lo() : !has_hi()) continue;
            const Msg &m = hm[1 + side]; BbxPeerArrays *dst = side == 0 ? lo : hi;
            if(!m.good){ failed = 1u; break; }
            dst->gc = m.gc;
            for(int k = 0; k < BBX_PEER_NPTR; k++){
                void *q = nullptr;
                if(cudaIpcOpenMemHandle(&q, m.h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess){ cudaGetLastError(); failed = 1u; break; }
                dst->p[k] = q; opened[n_opened++] = q;
            }
        }
        unsigned *dflag = (unsigned *)dm;
        BBX_CUC(cudaMemcpyAsync(dflag, &failed, sizeof(unsigned), cudaMemcpyHostToDevice, s));
        BBX_NCCL(g_nccl.AllReduce(dflag, dflag, 1, ncclUint32, ncclMax, comm, s));
        BBX_CUC(cudaMemcpyAsync(&failed, dflag, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        BBX_CUC(cudaStreamSynchronize(s));
        cudaFree(dm); cudaFreeHost(hm);
        *ok = failed ? 0 : 1;
        return 0;
    }
    int barrier() override { return 0; } // every collective above already orders the ranks
};

// --------------------------------------------------------------------------- in-process transport
struct LocalShared {
    int nranks = 0, attached = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0; unsigned long gen = 0; bool failed = false;
    struct Slot {
        BbxSeg send_lo[BBX_MAX_SEGS], send_hi[BBX_MAX_SEGS];
        int n_lo = 0, n_hi = 0;
        cudaEvent_t ready = nullptr, done = nullptr;
        int to_lo[BBX_NCOUNTS] = {0}, to_hi[BBX_NCOUNTS] = {0};
        unsigned red[16];
        BbxPeerArrays arrays; int want = 0;
    } slot[BBX_LOCAL_MAX_RANKS];
    // false on timeout / failure of a peer: the group is then poisoned (every later barrier fails)
    bool barrier(){
        std::unique_lock<std::mutex> lk(m);
        if(failed) return false;
        unsigned long my = gen;
        if(++arrived == nranks){ arrived = 0; gen++; cv.notify_all(); return true; }
        if(!cv.wait_for(lk, std::chrono::seconds(120), [&]{ return gen != my || failed; })){ failed = true; cv.notify_all(); return false; }
        return !failed;
    }
    void poison(){ std::lock_guard<std::mutex> lk(m); failed = true; cv.notify_all(); }
};
static std::mutex g_local_mutex;
static std::map<std::string, std::shared_ptr<LocalShared>> g_local_groups;

struct LocalComm : BbxComm {
    std::shared_ptr<LocalShared> sh;
    std::string key;
    unsigned *redhost = nullptr; // pinned
    int init(const char *name, int rank_, int nranks_){
        rank = rank_; nranks = nranks_; key = name ? name : "";
        if(nranks < 1 || nranks > BBX_LOCAL_MAX_RANKS || rank < 0 || rank >= nranks){ err = "bad rank / nranks"; return 1; }
        {
            std::lock_guard<std::mutex> lk(g_local_mutex);
            auto it = g_local_groups.find(key);
            if(it == g_local_groups.end() || it->second->attached >= it->second->nranks){
                sh = std::make_shared<LocalShared>(); sh->nranks = nranks; g_local_groups[key] = sh;
            }else sh = it->second;
            if(sh->nranks != nranks){ err = "nranks differs from the group's"; return 1; }
            sh->attached++;
        }
        BBX_CUC(cudaEventCreateWithFlags(&sh->slot[rank].ready, cudaEventDisableTiming));
        BBX_CUC(cudaEventCreateWithFlags(&sh->slot[rank].done, cudaEventDisableTiming));
        BBX_CUC(cudaMallocHost((void **)&redhost, 16 * sizeof(unsigned)));
        return 0;
    }
    ~LocalComm() override {
        if(sh){
            sh->poison(); // a member going away ends the group
            if(sh->slot[rank].ready) cudaEventDestroy(sh->slot[rank].ready);
            if(sh->slot[rank].done) cudaEventDestroy(sh->slot[rank].done);
            std::lock_guard<std::mutex> lk(g_local_mutex);
            auto it = g_local_groups.find(key);
            if(it != g_local_groups.end() && it->second == sh) g_local_groups.erase(it);
        }
        if(redhost) cudaFreeHost(redhost);
    }
    int fail(const char *what){ err = what; sh->poison(); return 1; }
    int exchange(cudaStream_t s, const BbxSeg *send_lo, const BbxSeg *recv_lo, int n_lo,
                 const BbxSeg *send_hi, const BbxSeg *recv_hi, int n_hi) override {
        LocalShared::Slot &me = sh->slot[rank];
        me.n_lo = n_lo; me.n_hi = n_hi;
        for(int k = 0; k < n_lo; k++) me.send_lo[k] = send_lo[k];
        for(int k = 0; k < n_hi; k++) me.send_hi[k] = send_hi[k];
        BBX_CUC(cudaEventRecord(me.ready, s));
        if(!sh->barrier()) return fail("local slab group: a peer did not reach the exchange");
        if(has_lo()){
            LocalShared::Slot &p = sh->slot[rank - 1];
            BBX_CUC(cudaStreamWaitEvent(s, p.ready, 0));
            if(p.n_hi != n_lo) return fail("local slab group: segment count mismatch (lo)");
            for(int k = 0; k < n_lo; k++){
                if(p.send_hi[k].bytes != recv_lo[k].bytes) return fail("local slab group: segment size mismatch (lo)");
                if(recv_lo[k].bytes) BBX_CUC(cudaMemcpyAsync(recv_lo[k].ptr, p.send_hi[k].ptr, recv_lo[k].bytes, cudaMemcpyDeviceToDevice, s));
            }
        }
        if(has_hi()){
            LocalShared::Slot &p = sh->slot[rank + 1];
            BBX_CUC(cudaStreamWaitEvent(s, p.ready, 0));
            if(p.n_lo != n_hi) return fail("local slab group: segment count mismatch (hi)");
            for(int k = 0; k < n_hi; k++){
                if(p.send_lo[k].bytes != recv_hi[k].bytes) return fail("local slab group: segment size mismatch (hi)");
                if(recv_hi[k].bytes) BBX_CUC(cudaMemcpyAsync(recv_hi[k].ptr, p.send_lo[k].ptr, recv_hi[k].bytes, cudaMemcpyDeviceToDevice, s));
            }
        }
        BBX_CUC(cudaEventRecord(me.done, s));
        if(!sh->barrier()) return fail("local slab group: a peer did not finish the exchange");
        // my send ranges may be overwritten only after the neighbours have copied them
        if(has_lo()) BBX_CUC(cudaStreamWaitEvent(s, sh->slot[rank - 1].done, 0));
        if(has_hi()) BBX_CUC(cudaStreamWaitEvent(s, sh->slot[rank + 1].done, 0));
        return 0;
    }
    int allreduce_max_u32(cudaStream_t s, unsigned *dev, int count) override {
        if(count > 16) return fail("allreduce_max_u32: count > 16");
        BBX_CUC(cudaMemcpyAsync(redhost, dev, count * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        BBX_CUC(cudaStreamSynchronize(s));
        memcpy(sh->slot[rank].red, redhost, count * sizeof(unsigned));
        if(!sh->barrier()) return fail("local slab group: allreduce");
        for(int r = 0; r < nranks; r++) for(int k = 0; k < count; k++) if(sh->slot[r].red[k] > redhost[k]) redhost[k] = sh->slot[r].red[k];
        if(!sh->barrier()) return fail("local slab group: allreduce");
        BBX_CUC(cudaMemcpyAsync(dev, redhost, count * sizeof(unsigned), cudaMemcpyHostToDevice, s));
        BBX_CUC(cudaStreamSynchronize(s)); // redhost is reused by the next call
        return 0;
    }
    int neighbor_counts(cudaStream_t s, const int *to_lo, const int *to_hi, int *from_lo, int *from_hi) override {
        (void)s;
        for(int k = 0; k < BBX_NCOUNTS; k++){ sh->slot[rank].to_lo[k] = to_lo[k]; sh->slot[rank].to_hi[k] = to_hi[k]; }
        if(!sh->barrier()) return fail("local slab group: counts");
        for(int k = 0; k < BBX_NCOUNTS; k++){
            from_lo[k] = has_lo() ? sh->slot[rank - 1].to_hi[k] : 0;
            from_hi[k] = has_hi() ? sh->slot[rank + 1].to_lo[k] : 0;
        }
        if(!sh->barrier()) return fail("local slab group: counts");
        return 0;
    }
    // one process, one device: the neighbours' pointers are usable as they are
    int share_arrays(cudaStream_t s, const BbxPeerArrays &mine, int want, BbxPeerArrays *lo, BbxPeerArrays *hi, int *ok) override {
        (void)s;
        memset(lo, 0, sizeof(*lo)); memset(hi, 0, sizeof(*hi));
        sh->slot[rank].arrays = mine; sh->slot[rank].want = want;
        if(!sh->barrier()) return fail("local slab group: share_arrays");
        int all = 1;
        for(int r = 0; r < nranks; r++) all &= sh->slot[r].want ? 1 : 0;
        if(has_lo()) *lo = sh->slot[rank - 1].arrays;
        if(has_hi()) *hi = sh->slot[rank + 1].arrays;
        if(!sh->barrier()) return fail("local slab group: share_arrays");
        *ok = all;
        return 0;
    }
    int barrier() override { return sh->barrier() ? 0 : fail("local slab group: barrier"); }
    bool shares_device() const override { return true; }
};
