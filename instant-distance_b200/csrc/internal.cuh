// internal.cuh — host-side index object, per-device shared context and kernel argument blocks (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/instant_distance_b200.h"
#include "hnsw_device.cuh"

namespace idb {

// Warps (= live queries / inserts) per CTA of the traversal kernels.  A CTA's resources are only handed to the next launch when ALL its
// warps have finished their last query, so fewer warps per CTA means the next batch moves in sooner at a batch boundary.
#ifndef IDB_WPC
#define IDB_WPC 4
#endif
constexpr int kSearchWarps = IDB_WPC;
constexpr int kSearchCtasPerSm = 16 / IDB_WPC;  // resident CTAs per SM -> 16 live queries per SM, <= 128 registers per thread
constexpr int kMaxCtasPerSm = 32 / IDB_WPC;     // upper bound for IDB_CTAS_PER_SM
constexpr int kRetryCtas = 32 / IDB_WPC;        // CTA slots of the (normally idle) overflow-retry pool -> 32 warps
constexpr int occ_for_warps(int warps_per_sm) { return warps_per_sm / IDB_WPC; }
constexpr uint32_t kRetrySlots = 1u << 18;  // hash slots per retry warp: 196k ids at 3/4 load (2M * ef <= 131k for M <= 64, ef <= 1024)
constexpr int kLanes = 4;              // submission lanes per index (own stream + per-call control state)

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
    TablePool pool;                    // per-warp scratch tables, claimed per CTA (hnsw_device.cuh)
    uint32_t gslots, gshift;           // big visited tier: words in use per warp / hash flavour: 32 - log2(gslots)
    uint32_t vis_mode;                 // flavour of the big visited tier (hnsw_device.cuh VisMode)
    uint32_t b16_cap_ids;              // b16 flavour: ids per query before the retry pass takes over
    uint32_t b16_nb;                   // b16 flavour: buckets in use over both segments
    uint64_t* out_keys;                // optional: nq x k packed (distance bits << 32 | id_map[pid]) for the sharded all-gather
    const uint32_t* id_map;            // optional: PointId -> caller's global row id
    int variant;                       // tuning variant of the kernel template (0 = default)
};

// Persisting-L2 access-policy window attached to a launch (the b16 visited tables), or none.
struct LaunchWindow {
    void* base = nullptr;
    size_t bytes = 0;
    float hit_ratio = 1.0f;
};

// ---------------------------------------------------------------------------------------------------------
// One per CUDA device, shared by every index on it (reference-counted): the pool of per-warp scratch tables (sized for the
// warps that can be RESIDENT, not per index or per call), the retry pool, and the device's persisting-L2 reservation.
// ---------------------------------------------------------------------------------------------------------
struct DeviceCtx {
    int device = 0;
    int num_sms = 148;
    int sm_ids = 148;                      // %nsmid: SM ids are < sm_ids, which can exceed the number of enabled SMs (148 of 160 on B200)
    int slots_per_sm = kSearchCtasPerSm;   // CTA slots per SM (IDB_CTAS_PER_SM)
    std::mutex mu;                         // held while tables are (re)allocated and while a launch that uses them is enqueued
    uint32_t* slot_masks = nullptr;        // sm_ids words + 1 (the retry pool)
    uint32_t n_tables = 0;                 // sm_ids * slots_per_sm * kSearchWarps (only those of enabled SMs are ever touched)
    uint32_t n_tables_live = 0;            // num_sms * slots_per_sm * kSearchWarps: how many can be in use at once
    // b16 tier: fixed stride per warp, a prefix of it in use per call
    uint32_t* b16_tables = nullptr;        // first segment of every table: what normal traversals use, under the persisting-L2 window
    uint32_t b16_stride = 0;               // u32 words per warp = b16_l2_bytes / 4
    uint32_t b16_l2_bytes = 32 * 1024;     // bytes per warp that keep all live tables inside the persisting part of L2
    uint32_t* b16_ext = nullptr;           // second segment (same size), used by traversals with a large ef; not under the window
    // atomic tiers (hash / bitmap): allocated on first use, regrown (device idle) when a call needs more
    uint32_t* big_tables = nullptr;
    uint32_t big_stride = 0;
    uint32_t* retry_tables = nullptr;      // kRetryCtas * kSearchWarps tables of kRetrySlots words
    uint64_t* tie_tables = nullptr;        // n_tables * kTieCap
    uint64_t* retry_ties = nullptr;        // kRetryCtas * kSearchWarps * kRetryTieCap
    int max_persist = 0, max_window = 0;
    size_t l2_reserved = 0;                // current cudaLimitPersistingL2CacheSize set by this library
    bool l2_allowed = true;                // idb_device_set_persisting_l2 / IDB_L2_PERSIST
    int refs = 0;

    static idb_status acquire(int device, DeviceCtx** out);
    static void release(DeviceCtx* c);
    idb_status ensure_big(uint32_t stride_words);          // caller holds mu
    idb_status reserve_l2(size_t bytes);                    // caller holds mu
    TablePool main_pool(bool b16) const;
    TablePool retry_pool() const;
    ~DeviceCtx();
};

// Per-call control state + host-API staging buffers; one per submission lane.  Calls on one lane are stream-ordered, so the
// buffers are reused without waiting; calls on different lanes overlap on the device.
struct Lane {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    unsigned char* ctrl = nullptr;   // [0..8) K1 work counter, [16..20) K1 fail count, [32..40) retry work counter, [48..52) retry fail count
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
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // host API: results land here first when the caller's output buffers are pageable (see HostOut in api.cu)
    unsigned char* h_out = nullptr; size_t h_out_cap = 0;   // pinned
    // asynchronous read-back of the control block of the lane's last call (how many queries overflowed the b16 tables)
    uint32_t* h_ctrl = nullptr;      // pinned, 32 words: [0,16) the sampled overflow tally, [16,32) the host API's read-back
    cudaEvent_t ev_ctrl = nullptr;
    bool ctrl_pending = false;
    int ctrl_b16 = 0;
    int last_b16 = 0;                // the lane's last call used the b16 visited flavour: 1 = first segment only, 2 = both segments
    void* win_base = nullptr;        // access-policy window currently attached to the stream (see Index::attach_window)
    size_t win_bytes = 0;
    uint32_t ctrl_ef = 0;
    uint64_t ctrl_nq = 0;
    uint64_t last_nq = 0;
    uint32_t last_launches = 0;
    void free_all();
};

// Device -> host copy of a batch's results through the lane's stream.  Output buffers in pinned (or registered) host memory receive
// the copies directly.  Pageable buffers do NOT: a device-to-pageable cudaMemcpyAsync blocks inside the driver until the copy has run
// (i.e. until this call's K1 has finished) and stalls the launches of other caller threads meanwhile — concurrent callers would never
// have a second batch queued behind the running one.  Those results are staged in the lane's pinned buffer and copied out after the
// stream has been synchronised.
struct HostOut {
    struct Part { void* user; const void* dev; size_t bytes; size_t off; };
    Part parts[3];
    int n = 0;
    bool staged = false;
    Lane* lane = nullptr;
    void add(void* user, const void* dev, size_t bytes) { if (user && bytes) parts[n++] = Part{user, dev, bytes, 0}; }
    cudaError_t enqueue(Lane& ln);   // after the kernels of the call
    void finish() const;             // after cudaStreamSynchronize
};

struct Index {
    int device = 0;
    int num_sms = 148;
    DeviceCtx* ctx = nullptr;
    Lane lanes[kLanes];
    cudaStream_t stream = nullptr;             // = lanes[0].stream: uploads, builds, the default lane
    std::mutex mu;                             // graph-level operations (build, export, id map)
    std::atomic<uint32_t> next_lane{0};        // host-API calls rotate over the lanes
    std::atomic<int> last_lane{0};

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
    bool rows_distinct = true;                 // no adjacency row lists a PointId twice (checked for adopted graphs)

    // tuning knobs (env IDB_OPT / IDB_VIS_MULT / IDB_VIS_TIER / IDB_B16_BYTES / IDB_VIS_SLOTS / IDB_VARIANT); none of them changes results
    uint32_t opt_flags = 0;       // L2 prefetch of rows/vectors: measured neutral-to-negative once 16 rows are in flight (profiles/r01_call4)
    uint32_t vis_mult = 4;        // hash flavour: slots = next_pow2(vis_mult * 2M * ef): load <= ~0.15, probe chains ~1
    uint32_t vis_slots_override = 0; // IDB_VIS_SLOTS (tests): exact hash-table size, to force the overflow -> retry path
    uint32_t b16_bytes_override = 0; // IDB_B16_BYTES (tests / sweeps): exact b16 table bytes per warp in use
    uint32_t b16_cap_16ths = 11;     // IDB_B16_CAP (tests): hand a query to the retry pass beyond this many sixteenths of the slots
    int vis_tier = -1;            // IDB_VIS_TIER: -1 auto (b16 when exact for this n, else bitmap / hash), 0 hash, 1 bitmap, 2 b16
    int variant = 0;              // IDB_VARIANT: alternative (rows in flight, CTAs/SM) instantiations of K1
    // Adaptive: when more than 1 in 1000 traversals of a call overflowed the b16 tables (data whose traversals visit more ids than
    // the tables were sized for), later calls with that ef or a larger one use the DRAM-resident atomic flavours instead of paying
    // for the retry pass.  Results are identical either way.
    std::atomic<uint32_t> b16_demote_ef[2] = {{0xFFFFFFFFu}, {0xFFFFFFFFu}};  // [0] first-segment tables, [1] two-segment tables
    int b16_level = 0;            // set by select_visited_tier (under ctx->mu): which b16 size the selected tier is (0 = not b16)
    void note_overflows(uint32_t ef, uint64_t n_work, uint32_t overflowed, int level);
    bool profiling = false;

    ~Index();
    idb_status init_device(int dev);
    idb_status upload(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef, const uint32_t* zero,
                      uint32_t n_upper, const uint32_t* const* upper, const uint64_t* upper_n);
    GraphView view() const;
    idb_status narrow_points_to_bf16();                                  // d_points (f32) -> d_points_bf16, frees d_points
    idb_status copy_points_f32(float* host_out, uint64_t r0, uint64_t m);  // rows [r0, r0+m) as n x dim f32 on the host
    int search_grid() const;
    // Fills the visited-tier fields of `a` (pool, gslots, ...) for a traversal with this ef and returns the launch window.
    // Caller holds ctx->mu.
    idb_status select_visited_tier(uint32_t ef, SearchArgs& a, LaunchWindow& win);
    // The persisting-L2 window on the b16 tables rides on every launch as a launch attribute; it is ALSO kept as a stream attribute,
    // because profilers that replay a kernel (ncu) re-launch it without its launch attributes.
    idb_status attach_window(Lane& ln, const LaunchWindow& win);
    idb_status ensure_lane_scratch(Lane& ln, uint64_t nq);
    idb_status enqueue_search(Lane& ln, const float* d_queries_padded, uint64_t nq, uint32_t ef, uint32_t k, uint32_t* d_ids,
                              float* d_dist, uint32_t* d_len, uint64_t* out_keys);
    Lane& pick_lane();
};

cudaError_t fill_u32(uint32_t* p, size_t n, uint32_t v, cudaStream_t st);
cudaError_t ensure_u32(uint32_t*& p, size_t& cap, size_t need);
cudaError_t ensure_u64(uint64_t*& p, size_t& cap, size_t need);
cudaError_t ensure_f32(float*& p, size_t& cap, size_t need);

}  // namespace idb
