// peer_exchange.cu — the exchange step of the row-sharded searches over NVLink peer
// memory instead of a collective library call: every rank STORES its packed top-k record straight into a
// slot of every peer's gather buffer (P2P stores through NVLink / NVSwitch), publishes a flag, waits for
// the peers' flags on its own memory and runs the K7 select — for a single query all of it in ONE
// kernel (push + wait + rank merge), so a sharded step is scan -> unpack -> exchange_merge with no
// NCCL launch, no stream hand-off and no separate merge launch. Records of many queries (a 1024-query
// batch is 1.6 MB) take two launches: a multi-CTA push (which never waits, so it cannot deadlock against
// a peer's waiting CTAs) and the per-query wait + merge.
//
// Buffers: each rank owns gather[2][world][record_bytes] + flags[2][world]; parity = epoch & 1. A peer can
// run at most one step ahead (its wait for step s+1 needs my flag s+1, which my stream orders after my
// merge of step s), so two parities suffice. Flags carry the epoch number; data is fenced system-wide
// before the flag store and read back with volatile loads (the writer was another GPU: L1 may be stale).
//
// The buffers are mapped across processes with CUDA IPC handles (one process per GPU under torchrun; the
// handles travel over the process group once, at set-up) or addressed directly inside one process.
#include "peer_exchange.h"

#include <vector>

#include "runtime.h"
#include "topk.cuh"

namespace vb {

namespace {

constexpr unsigned long long kPeerTimeoutNs = 4ull * 1000 * 1000 * 1000;   // a dead peer must not hang the GPU

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_vol_u64(const void* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_vol_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

struct PeerPtrs {
    unsigned char* gather[kMaxPeers];   // every rank's gather buffer as addressable from this rank
    uint32_t* flags[kMaxPeers];         // every rank's flag array
};

struct RecordLayout {
    uint32_t nq, k_in, k_out;
    uint32_t off_keys, off_values, off_rows, off_counts, bytes;   // packed record (sharded.py packed_layout)
};

// Copies this rank's record into slot [parity][rank] of every peer (16-byte words, all threads of the grid).
__device__ __forceinline__ void push_record(const PeerPtrs& pp, const unsigned char* record, uint32_t bytes, uint32_t world,
                                            uint32_t rank, uint32_t parity, uint32_t tid, uint32_t nthreads) {
    const uint32_t words = bytes >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(record);
    for (uint32_t i = tid; i < words * world; i += nthreads) {
        const uint32_t peer = i / words, w = i - peer * words;
        uint4* dst = reinterpret_cast<uint4*>(pp.gather[peer] + ((size_t)parity * world + rank) * bytes);
        dst[w] = src[w];
    }
}

// Waits (threads 0..world-1, one peer each) until every rank's flag of this parity carries `epoch`.
__device__ __forceinline__ bool wait_flags(const uint32_t* my_flags, uint32_t world, uint32_t parity, uint32_t epoch) {
    bool ok = true;
    if (threadIdx.x < world) {
        const uint32_t* f = my_flags + parity * kMaxPeers + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) != epoch) {
            if (global_timer_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }
            __nanosleep(64);
        }
    }
    return __syncthreads_and(ok);
}

// K7 over the `world` records gathered in this rank's own buffer (select.cu's topk_merge_kernel with
// volatile loads), one CTA per query.
__device__ __forceinline__ void merge_gathered(unsigned char* smem, const unsigned char* gathered, const RecordLayout& L,
                                               uint32_t world, uint32_t cap, uint32_t qi, u64* keys_out, float* values_out,
                                               u64* rows_out, uint32_t* counts_out) {
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    Collector col;
    col.init(smem, &s_thresh, &s_count, cap, L.k_out);
    __syncthreads();
    auto rec = [&](uint32_t l) { return gathered + (size_t)l * L.bytes; };
    const uint32_t k_in = L.k_in, k_out = L.k_out;
    for (uint32_t l = threadIdx.x; l < world; l += blockDim.x) {
        const uint32_t cnt = min(ld_vol_u32(rec(l) + L.off_counts + 4 * qi), k_in);
        if (cnt >= k_out) {
            const u64 kth = ld_vol_u64(rec(l) + L.off_keys + 8 * ((size_t)qi * k_in + k_out - 1));
            if (kth != kKeyMax) atomicMin(col.thresh, kth + 1);
        }
    }
    __syncthreads();
    collector_merge_lists(
        col, world, k_in, [&](uint32_t l) { return ld_vol_u32(rec(l) + L.off_counts + 4 * qi); },
        [&](uint32_t l, uint32_t i) { return ld_vol_u64(rec(l) + L.off_keys + 8 * ((size_t)qi * k_in + i)); },
        [&](uint32_t l, uint32_t i) { return ((u64)l << 32) | i; });
    const uint32_t total = *col.count;
    for (uint32_t i = threadIdx.x; i < k_out; i += blockDim.x) {
        const size_t o = (size_t)qi * k_out + i;
        if (i < total) {
            const u64 pos = col.pays[i];
            const uint32_t l = (uint32_t)(pos >> 32);
            const size_t src = (size_t)qi * k_in + (uint32_t)pos;
            keys_out[o] = col.keys[i];
            if (values_out) values_out[o] = __uint_as_float(ld_vol_u32(rec(l) + L.off_values + 4 * src));
            if (rows_out) rows_out[o] = ((u64)l << 32) | ld_vol_u32(rec(l) + L.off_rows + 4 * src);
        } else {
            keys_out[o] = kKeyMax;
            if (values_out) values_out[o] = 0.0f;
            if (rows_out) rows_out[o] = 0;
        }
    }
    if (threadIdx.x == 0) counts_out[qi] = total;
}

// Large records, launch 1: every CTA copies its share; the last one to finish publishes the flags.
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerPtrs pp, const unsigned char* record, uint32_t bytes, uint32_t world, uint32_t rank, uint32_t epoch,
                 uint32_t* ticket) {
    __shared__ uint32_t s_last;
    const uint32_t parity = epoch & 1u;
    push_record(pp, record, bytes, world, rank, parity, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < world) st_release_sys(pp.flags[threadIdx.x] + parity * kMaxPeers + rank, epoch);
    if (threadIdx.x == 0) *ticket = 0u;
}

// Large records, launch 2: one CTA per query waits for the flags, then merges its query.
__global__ void __launch_bounds__(256)
peer_wait_merge_kernel(PeerPtrs pp, RecordLayout L, uint32_t world, uint32_t rank, uint32_t epoch, uint32_t cap,
                       u64* keys_out, float* values_out, u64* rows_out, uint32_t* counts_out, uint32_t* error) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t parity = epoch & 1u, qi = blockIdx.x;
    if (!wait_flags(pp.flags[rank], world, parity, epoch)) {
        if (threadIdx.x == 0) { *error = 1u; counts_out[qi] = 0u; }
        return;
    }
    merge_gathered(smem, pp.gather[rank] + (size_t)parity * world * L.bytes, L, world, cap, qi, keys_out, values_out,
                   rows_out, counts_out);
}

// ---- one query (or a handful): push + flag + wait + RANK MERGE in one launch ---------------------------------
// The lists are sorted and keys are unique (the low word is the global id rank), so an entry's position in the
// merged order is simply its own index plus, for every other list, the number of entries below it (a binary
// search): no sort, no atomics, every entry independent. All lists are staged in shared memory (world x k_in x
// 8 B <= 64 KB). k = 10 from 8 shards is one 128-thread CTA; the 1000 Hamming candidates of quantized_search x 8
// shards (a single CTA pushing 8000 entries through a sorting collector cost ~100 us) are split over 8 CTAs: their
// pushes never wait, the last one to finish publishes the flags, then every CTA waits and ranks its share.
constexpr uint32_t kRankThreads = 512;
constexpr uint32_t kRankCtas = 8;

__global__ void __launch_bounds__(kRankThreads)
peer_exchange_rank_merge_kernel(PeerPtrs pp, const unsigned char* record, RecordLayout L, uint32_t world, uint32_t rank,
                                uint32_t epoch, u64* keys_out, float* values_out, u64* rows_out, uint32_t* counts_out,
                                uint32_t* ticket, uint32_t* error) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_cnt[kMaxPeers];
    __shared__ uint32_t s_last;
    u64* s_keys = reinterpret_cast<u64*>(smem);                 // [world][k_in]
    const uint32_t parity = epoch & 1u;
    push_record(pp, record, L.bytes, world, rank, parity, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < world) st_release_sys(pp.flags[threadIdx.x] + parity * kMaxPeers + rank, epoch);
        if (threadIdx.x == 0) *ticket = 0u;
    }
    if (!wait_flags(pp.flags[rank], world, parity, epoch)) {
        if (threadIdx.x == 0 && blockIdx.x == 0) { *error = 1u; for (uint32_t q = 0; q < L.nq; ++q) counts_out[q] = 0u; }
        return;
    }
    const unsigned char* gathered = pp.gather[rank] + (size_t)parity * world * L.bytes;
    const uint32_t k_in = L.k_in, k_out = L.k_out;
    for (uint32_t qi = 0; qi < L.nq; ++qi) {
        __syncthreads();
        if (threadIdx.x < world)
            s_cnt[threadIdx.x] = min(ld_vol_u32(gathered + (size_t)threadIdx.x * L.bytes + L.off_counts + 4 * qi), k_in);
        for (uint32_t s = threadIdx.x; s < world * k_in; s += blockDim.x) {
            const uint32_t l = s / k_in, i = s - l * k_in;
            s_keys[s] = ld_vol_u64(gathered + (size_t)l * L.bytes + L.off_keys + 8 * ((size_t)qi * k_in + i));
        }
        __syncthreads();
        uint32_t total = 0;
        for (uint32_t l = 0; l < world; ++l) total += s_cnt[l];
        const uint32_t kept = min(total, k_out);
        for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < world * k_in; s += gridDim.x * blockDim.x) {
            const uint32_t l = s / k_in, i = s - l * k_in;
            if (i >= s_cnt[l] || i >= k_out) continue;
            const u64 key = s_keys[s];
            uint32_t pos = i;
            for (uint32_t m = 0; m < world && pos < k_out; ++m) {
                if (m == l) continue;
                const u64* lk = s_keys + m * k_in;
                uint32_t lo = 0, hi = s_cnt[m];
                while (lo < hi) {                       // entries of list m below `key` (keys are unique)
                    const uint32_t mid = (lo + hi) >> 1;
                    if (lk[mid] < key) lo = mid + 1; else hi = mid;
                }
                pos += lo;
            }
            if (pos >= k_out) continue;
            const size_t o = (size_t)qi * k_out + pos, src = (size_t)qi * k_in + i;
            const unsigned char* rec = gathered + (size_t)l * L.bytes;
            keys_out[o] = key;
            if (values_out) values_out[o] = __uint_as_float(ld_vol_u32(rec + L.off_values + 4 * src));
            if (rows_out) rows_out[o] = ((u64)l << 32) | ld_vol_u32(rec + L.off_rows + 4 * src);
        }
        if (blockIdx.x == 0) {
            for (uint32_t i = kept + threadIdx.x; i < k_out; i += blockDim.x) {
                const size_t o = (size_t)qi * k_out + i;
                keys_out[o] = kKeyMax;
                if (values_out) values_out[o] = 0.0f;
                if (rows_out) rows_out[o] = 0;
            }
            if (threadIdx.x == 0) counts_out[qi] = kept;
        }
    }
}

}  // namespace

struct PeerExchange::Impl {
    PeerPtrs pp{};
    std::vector<void*> ipc_opened;
    uint32_t* d_ticket = nullptr;   // [0] push ticket, [1] sticky error
};

PeerExchange::PeerExchange(int world, int rank, size_t record_bytes, int device)
    : world_(world), rank_(rank), device_(device), record_bytes_(record_bytes), impl_(new Impl()) {}

PeerExchange::~PeerExchange() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (void* p : impl_->ipc_opened) cudaIpcCloseMemHandle(p);
    if (buf_) cudaFree(buf_);
    if (impl_->d_ticket) cudaFree(impl_->d_ticket);
    delete impl_;
}

size_t PeerExchange::buffer_bytes() const { return 2 * (size_t)world_ * record_bytes_ + 2 * kMaxPeers * sizeof(uint32_t) + 256; }
size_t PeerExchange::flags_offset() const { return (2 * (size_t)world_ * record_bytes_ + 255) & ~(size_t)255; }

Status PeerExchange::allocate() {
    if (world_ < 1 || world_ > kMaxPeers || rank_ < 0 || rank_ >= world_) return Status::Cuda("peer exchange: bad world / rank");
    if (record_bytes_ == 0 || (record_bytes_ & 15)) return Status::Cuda("peer exchange: record size must be a multiple of 16");
    VB_CUDA(cudaSetDevice(device_));
    const size_t bytes = flags_offset() + 2 * kMaxPeers * sizeof(uint32_t);
    VB_CUDA(cudaMalloc(&buf_, bytes));
    VB_CUDA(cudaMemset(buf_, 0, bytes));
    VB_CUDA(cudaMalloc(&impl_->d_ticket, 64));
    VB_CUDA(cudaMemset(impl_->d_ticket, 0, 64));
    VB_CUDA(cudaDeviceSynchronize());
    return Status::Ok();
}

Status PeerExchange::export_handle(unsigned char out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    VB_CUDA(cudaSetDevice(device_));
    VB_CUDA(cudaIpcGetMemHandle(&h, buf_));
    std::memcpy(out, &h, 64);
    return Status::Ok();
}

void PeerExchange::set_peer(int r, unsigned char* base) {
    impl_->pp.gather[r] = base;
    impl_->pp.flags[r] = reinterpret_cast<uint32_t*>(base + flags_offset());
}

Status PeerExchange::connect_ipc(const unsigned char* handles) {
    VB_CUDA(cudaSetDevice(device_));
    for (int r = 0; r < world_; ++r) {
        if (r == rank_) { set_peer(r, static_cast<unsigned char*>(buf_)); continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        VB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        impl_->ipc_opened.push_back(p);
        set_peer(r, static_cast<unsigned char*>(p));
    }
    connected_ = true;
    return Status::Ok();
}

Status PeerExchange::connect_local(PeerExchange* const* peers) {
    VB_CUDA(cudaSetDevice(device_));
    for (int r = 0; r < world_; ++r) {
        PeerExchange* p = peers[r];
        if (p->world_ != world_ || p->record_bytes_ != record_bytes_ || p->rank_ != r || !p->buf_)
            return Status::Cuda("peer exchange: mismatched local peers");
        if (p->device_ != device_) {
            int can = 0;
            VB_CUDA(cudaDeviceCanAccessPeer(&can, device_, p->device_));
            if (!can) return Status::Cuda("peer exchange: no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(p->device_, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return Status::Cuda(cudaGetErrorString(e));
            cudaGetLastError();
        }
        set_peer(r, static_cast<unsigned char*>(p->buf_));
    }
    connected_ = true;
    return Status::Ok();
}

static Status check_layout(const PeerExchange& px, const PeerRecord& r) {
    if (r.nq == 0 || r.k_in == 0 || r.k_out == 0) return Status::Cuda("peer exchange: empty record");
    if (r.k_in > (uint32_t)kMaxFusedK || r.k_out > (uint32_t)kMaxFusedK) return Status::Cuda("peer exchange: k beyond 1024");
    if (r.bytes != px.record_bytes()) return Status::Cuda("peer exchange: record size differs from the buffers'");
    return Status::Ok();
}

static RecordLayout to_layout(const PeerRecord& r) {
    RecordLayout L;
    L.nq = r.nq; L.k_in = r.k_in; L.k_out = r.k_out;
    L.off_keys = r.off_keys; L.off_values = r.off_values; L.off_rows = r.off_rows; L.off_counts = r.off_counts;
    L.bytes = (uint32_t)r.bytes;
    return L;
}

static uint32_t merge_cap(const PeerRecord& r) {
    uint32_t cap = 256;
    while (cap < 2 * r.k_out || cap < r.k_out + r.k_in) cap <<= 1;
    return cap;
}

Status PeerExchange::push(const void* d_record, const PeerRecord& rec, cudaStream_t stream) {
    if (!connected_) return Status::Cuda("peer exchange: not connected");
    VB_TRY(check_layout(*this, rec));
    ++epoch_;
    const uint32_t words = (uint32_t)(rec.bytes >> 4) * (uint32_t)world_;
    const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(64, (words + 1023) / 1024));
    peer_push_kernel<<<grid, 256, 0, stream>>>(impl_->pp, static_cast<const unsigned char*>(d_record), (uint32_t)rec.bytes,
                                               (uint32_t)world_, (uint32_t)rank_, epoch_, impl_->d_ticket);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

Status PeerExchange::wait_merge(const PeerRecord& rec, u64* d_keys_out, float* d_values_out, u64* d_rows_out,
                                uint32_t* d_counts_out, cudaStream_t stream) {
    if (!connected_) return Status::Cuda("peer exchange: not connected");
    VB_TRY(check_layout(*this, rec));
    const uint32_t cap = merge_cap(rec);
    const size_t smem = (size_t)cap * 16;
    VB_TRY(ensure_dynamic_smem_for(peer_wait_merge_kernel, smem));
    peer_wait_merge_kernel<<<rec.nq, 256, smem, stream>>>(impl_->pp, to_layout(rec), (uint32_t)world_, (uint32_t)rank_, epoch_,
                                                          cap, d_keys_out, d_values_out, d_rows_out, d_counts_out,
                                                          impl_->d_ticket + 1);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

Status PeerExchange::exchange_merge(const void* d_record, const PeerRecord& rec, u64* d_keys_out, float* d_values_out,
                                    u64* d_rows_out, uint32_t* d_counts_out, cudaStream_t stream) {
    if (!connected_) return Status::Cuda("peer exchange: not connected");
    VB_TRY(check_layout(*this, rec));
    VB_CUDA(cudaSetDevice(device_));
    if (rec.nq > 4) {   // many queries: multi-CTA push, then one waiting CTA per query
        VB_TRY(push(d_record, rec, stream));
        return wait_merge(rec, d_keys_out, d_values_out, d_rows_out, d_counts_out, stream);
    }
    ++epoch_;
    // one fused launch: push + flag + wait + rank merge. Short lists (k = 10 from 8 shards) need one small CTA; long
    // ones (1000 candidates x 8) are ranked by 8 CTAs.
    const bool many = (size_t)world_ * rec.k_in > 1024;
    const size_t smem_r = (size_t)world_ * rec.k_in * sizeof(u64);
    VB_TRY(ensure_dynamic_smem_for(peer_exchange_rank_merge_kernel, smem_r));
    peer_exchange_rank_merge_kernel<<<many ? kRankCtas : 1, many ? kRankThreads : 128, smem_r, stream>>>(
        impl_->pp, static_cast<const unsigned char*>(d_record), to_layout(rec), (uint32_t)world_, (uint32_t)rank_, epoch_,
        d_keys_out, d_values_out, d_rows_out, d_counts_out, impl_->d_ticket, impl_->d_ticket + 1);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

Status PeerExchange::error_state(uint32_t* out) {
    VB_CUDA(cudaSetDevice(device_));
    VB_CUDA(cudaMemcpy(out, impl_->d_ticket + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return Status::Ok();
}

}  // namespace vb
