// internal.cuh — host-side index object and kernel argument blocks (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <new>
#include <vector>

#include "../../include/instant_distance_b200.h"
#include "hnsw_device.cuh"

namespace idb {

constexpr int kSearchWarps = 4;        // warps (= live queries) per CTA
constexpr int kSearchCtasPerSm = 4;    // resident CTAs per SM -> 16 live queries per SM, <= 128 registers per thread
constexpr int kMaxCtasPerSm = 8;       // upper bound over the tuning variants (scratch is sized for it)
constexpr int kRetryWarps = 32;        // warps of the (normally idle) overflow-retry pass
constexpr uint32_t kRetrySlots = 1u << 21;

extern thread_local char g_err[512];
idb_status fail(idb_status st, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (expr);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return ::idb::fail(e__ == cudaErrorMemoryAllocation ? IDB_ERR_OOM : IDB_ERR_CUDA, "CUDA error %s at %s:%d (%s)", \
                               cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__));             \
    } while (0)

struct SearchArgs {
    GraphView g;
    const float4* queries;             // nq x nchunks float4 (zero padded rows)
    unsigned long long n_work;         // number of work items ...
    const uint32_t* n_work_dev;        // ... or, if non-null, read it from device memory (retry pass)
    const uint32_t* work_list;         // optional indirection: work item -> query index
    uint32_t ef, k;
    uint32_t* out_ids;
    float* out_dist;
    uint32_t* out_len;
    uint32_t* counters;                // nq x 4 u32 or null
    uint32_t* status;                  // nq
    unsigned long long* work_counter;
    uint32_t* fail_count;
    uint32_t* fail_list;               // may be null (retry pass)
    uint32_t* vis_tables;
    uint32_t gslots, gshift;           // words in use per warp / 32 - log2(gslots) (hash and bucket flavours: gslots is a power of two)
    uint32_t vis_stride;               // words between consecutive warps' tables (>= gslots)
    uint32_t vis_mode;                 // flavour of the big visited tier (hnsw_device.cuh VisMode)
    uint64_t* tie_tables;
    uint64_t* out_keys;                // optional: nq x k packed (distance bits << 32 | id_map[pid]) for the sharded all-gather
    const uint32_t* id_map;            // optional: PointId -> caller's global row id
    int variant;                       // tuning variant of the kernel template (0 = default)
};

struct Scratch {
    uint32_t* vis_tables = nullptr;
    uint32_t gslots = 0;          // hash slots per warp (power of two)
    uint32_t bm_words = 0;        // != 0: K1 may use the bitmap flavour with this many words per warp
    uint32_t* bucket_tables = nullptr;  // bucket-set flavour: compact tables (bucket_slots words per warp) under a persisting-L2 window
    uint32_t bucket_slots = 0;    // != 0: K1 uses the bucket-set flavour
    uint32_t bucket_cap = 0;      // slots per warp the buffer was allocated for
    uint32_t vis_stride = 0;      // words per warp actually allocated (>= gslots; >= bitmap words when the bitmap flavour is on)
    uint32_t* retry_tables = nullptr;
    uint64_t* tie_tables = nullptr;
    unsigned char* ctrl = nullptr;
    uint32_t* status = nullptr;   size_t status_cap = 0;
    uint32_t* fail_list = nullptr; size_t fail_cap = 0;
    uint32_t* counters = nullptr; size_t counters_cap = 0;
    float* q = nullptr;           size_t q_cap = 0;
    uint32_t* ids = nullptr;      size_t ids_cap = 0;
    float* dist = nullptr;        size_t dist_cap = 0;
    uint32_t* len = nullptr;      size_t len_cap = 0;
    // sharded search
    uint64_t* keys_local = nullptr; size_t keys_local_cap = 0;
    uint64_t* keys_all = nullptr;   size_t keys_all_cap = 0;
    float* q2 = nullptr;          size_t q2_cap = 0;
    uint32_t* ids2 = nullptr;     size_t ids2_cap = 0;
};

struct Index {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;

    uint64_t n = 0;
    uint32_t dim = 0, nchunks = 0, M = 32, ef_search = 100;
    float* d_points = nullptr;                 // n x nchunks*4 f32 (PointId order); null when the rows are stored as bf16
    uint16_t* d_points_bf16 = nullptr;         // n x nchunks*4 bf16 (storage = IDB_STORAGE_BF16)
    bool bf16 = false;
    uint32_t* d_zero = nullptr;                // n x 2M
    std::vector<uint32_t*> d_upper;            // [l-1] -> n_l x M
    std::vector<uint64_t> upper_n;
    const uint32_t** d_upper_ptrs = nullptr;   // device copy of the pointer table
    uint32_t* d_id_map = nullptr;              // shard: PointId -> global row id (idb_index_set_id_map)
    uint64_t* pending_out_keys = nullptr;      // set by the sharded path around enqueue_search

    Scratch sc;
    uint64_t last_nq = 0;
    // tuning knobs (env IDB_OPT / IDB_VIS_MULT / IDB_VIS_BUCKETS / IDB_VIS_BITMAP / IDB_CTAS_PER_SM); none of them changes results
    uint32_t opt_flags = 0;       // L2 prefetch of rows/vectors: measured neutral-to-negative once 16 rows are in flight (profiles/r01_call4)
    uint32_t vis_mult = 4;        // visited table slots = next_pow2(vis_mult * 2M * ef): load <= ~0.15, probe chains ~1
    int ctas_per_sm = kSearchCtasPerSm;
    uint32_t vis_slots_override = 0; // IDB_VIS_SLOTS (tests): exact per-warp visited-table size, to force the overflow -> retry path
    int vis_bitmap = 1;           // IDB_VIS_BITMAP: 0 hash set, 1 bitmap over PointIds when it is no bigger than 2x the hash table
    uint32_t bucket_slots_override = 0; // IDB_BUCKET_SLOTS (tests): exact per-warp bucket-set size
    int vis_buckets = 1;          // IDB_VIS_BUCKETS: 1 = bucket set when the tables of all resident warps fit the persisting part of L2
    int variant = 0;              // IDB_VARIANT: alternative (rows in flight, CTAs/SM) instantiations of K1
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t last_launches = 0;

    ~Index();
    idb_status init_device(int dev);
    idb_status upload(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef, const uint32_t* zero,
                      uint32_t n_upper, const uint32_t* const* upper, const uint64_t* upper_n);
    GraphView view() const;
    idb_status narrow_points_to_bf16();                                  // d_points (f32) -> d_points_bf16, frees d_points
    idb_status copy_points_f32(float* host_out, uint64_t r0, uint64_t m);  // rows [r0, r0+m) as n x dim f32 on the host
    int search_grid() const;
    idb_status ensure_search_scratch(uint32_t ef, uint64_t nq, uint32_t k);
    idb_status enqueue_search(const float* d_queries_padded, uint64_t nq, uint32_t ef, uint32_t k, uint32_t* d_ids, float* d_dist,
                              uint32_t* d_len);
};

cudaError_t fill_u32(uint32_t* p, size_t n, uint32_t v, cudaStream_t st);
cudaError_t ensure_u32(uint32_t*& p, size_t& cap, size_t need);
cudaError_t ensure_u64(uint64_t*& p, size_t& cap, size_t need);
cudaError_t ensure_f32(float*& p, size_t& cap, size_t need);

}  // namespace idb
