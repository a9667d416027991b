// peer_exchange.h — top-k record exchange over NVLink peer memory + fused select (see peer_exchange.cu).
#pragma once
#include "common.cuh"

namespace vb {

constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch box

// One shard's packed record (the layout sharded.py's packed_layout / vb_flat_search_device write):
// keys[nq][k_in] u64 | values[nq][k_in] f32 | rows[nq][k_in] u32 | counts[nq] u32, `bytes` in all (16-byte multiple).
struct PeerRecord {
    uint32_t nq = 1, k_in = 0, k_out = 0;
    uint32_t off_keys = 0, off_values = 0, off_rows = 0, off_counts = 0;
    size_t bytes = 0;
};

class PeerExchange {
  public:
    PeerExchange(int world, int rank, size_t record_bytes, int device);
    ~PeerExchange();
    Status allocate();
    // One process per GPU: the buffer's CUDA IPC handle goes to every peer, which maps it.
    Status export_handle(unsigned char out[64]);
    Status connect_ipc(const unsigned char* handles /* [world][64], own slot ignored */);
    // One process, several GPUs (or several shards on one GPU in tests): direct pointers.
    Status connect_local(PeerExchange* const* peers /* [world], ordered by rank */);

    // push + flag + wait + K7 merge of the world's records; outputs as vb_topk_merge_device
    // (rows_out = shard << 32 | row). Stream-ordered; every rank must make the same sequence of calls.
    Status exchange_merge(const void* d_record, const PeerRecord& rec, u64* d_keys_out, float* d_values_out,
                          u64* d_rows_out, uint32_t* d_counts_out, cudaStream_t stream);
    // The two halves, for callers that drive several ranks from one thread (the push of EVERY rank must be
    // enqueued before any rank's wait when the waiting grid could fill the device).
    Status push(const void* d_record, const PeerRecord& rec, cudaStream_t stream);
    Status wait_merge(const PeerRecord& rec, u64* d_keys_out, float* d_values_out, u64* d_rows_out,
                      uint32_t* d_counts_out, cudaStream_t stream);
    // 1 when a wait timed out (a peer never published): results of that step are empty. Synchronises.
    Status error_state(uint32_t* out);

    size_t record_bytes() const { return record_bytes_; }
    int world() const { return world_; }
    int rank() const { return rank_; }

  private:
    struct Impl;
    size_t buffer_bytes() const;
    size_t flags_offset() const;
    void set_peer(int r, unsigned char* base);
    int world_, rank_, device_;
    size_t record_bytes_;
    void* buf_ = nullptr;
    uint32_t epoch_ = 0;
    bool connected_ = false;
    Impl* impl_;
};

}  // namespace vb

struct vb_peer {
    vb::PeerExchange* impl;
};
