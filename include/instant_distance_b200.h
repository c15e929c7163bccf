/*
 * instant_distance_b200.h — C ABI of the B200-native HNSW build-and-search engine.
 *
 * This is the drop-in boundary for djc/instant-distance's f32-vector hot path.  Every entry point below
 * names the reference interface it replaces (file:line under the reference tree; core =
 * instant-distance/src/lib.rs, types = instant-distance/src/types.rs, py = instant-distance-py/src/lib.rs).
 * A Rust `-sys` binding, the PyO3 module, or any other FFI binds exactly these symbols.
 *
 * Conventions
 *   - Plain C types only; all index state lives in GPU HBM behind an opaque handle.
 *   - Host-buffer calls copy in/out; the caller keeps ownership of every buffer it passes.
 *   - Every function returns an idb_status; idb_last_error() gives the thread-local message.
 *   - There is NO CPU fallback: without a CUDA device every compute call returns IDB_ERR_CUDA.
 *   - Points are f32 vectors under squared-L2 (the reference's FloatArray metric, py:378-421), any dim >= 1,
 *     computed in one canonical fp32 summation order so results are bit-reproducible.
 *   - PointIds are u32; IDB_INVALID (u32::MAX) is the reference's INVALID sentinel (types:293).
 */
#ifndef INSTANT_DISTANCE_B200_H
#define INSTANT_DISTANCE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDB_INVALID 0xFFFFFFFFu
#define IDB_STORAGE_F32 0u
#define IDB_STORAGE_BF16 1u

#if defined(__GNUC__)
#define IDB_API __attribute__((visibility("default")))
#else
#define IDB_API
#endif

typedef enum idb_status {
    IDB_OK = 0,
    IDB_ERR_INVALID_ARG = 1,   /* null pointer, dim == 0, N >= u32::MAX (core:256), unsupported M / ef ... */
    IDB_ERR_OOM = 2,           /* host or device allocation failed */
    IDB_ERR_CUDA = 3,          /* no device / CUDA runtime error (message has the CUDA error string) */
    IDB_ERR_NCCL = 4,
    IDB_ERR_IO = 5,
    IDB_ERR_FORMAT = 6,        /* malformed index file */
    IDB_ERR_CAPACITY = 7,      /* an internal per-query structure overflowed even after the retry pass */
    IDB_ERR_UNSUPPORTED = 8
} idb_status;

/* Opaque index handle: replaces `Hnsw<P>` (core:193-199) for P = f32 vector. */
typedef struct idb_index idb_index;

/* Builder (core:23-31) + Heuristic (core:115-119).  M is a compile-time const 32 in the reference
 * (core:787); it is a run-time field here because BASELINE.json's configs name M = 16 and M = 24. */
typedef struct idb_params {
    uint32_t M;                 /* reference: const M = 32 (core:787); supported 2..64 */
    uint32_t ef_construction;   /* Builder::ef_construction (core:35-38), default 100 (core:105) */
    uint32_t ef_search;         /* Builder::ef_search       (core:44-47), default 100 (core:104) */
    float    ml;                /* Builder::ml              (core:57-60), default 1/ln(M) (core:107) */
    uint64_t seed;              /* Builder::seed            (core:65-68) */
    int32_t  heuristic;         /* Builder::select_heuristic(Some/None) (core:49-52); default Some */
    int32_t  extend_candidates; /* Heuristic::extend_candidates (core:117), default false */
    int32_t  keep_pruned;       /* Heuristic::keep_pruned       (core:118), default true  */
    uint32_t insert_batch;      /* GPU build: concurrent inserts per step (rayon's worker count in the
                                   reference, core:316-318).  0 = auto, 1 = strictly sequential order. */
    int32_t  device;            /* CUDA device ordinal */
    uint32_t storage;           /* IDB_STORAGE_F32 (default) or IDB_STORAGE_BF16: rows rounded to bf16 (RNE) and kept in HBM at
                                   half the bytes; distances still accumulate in fp32 in the same canonical order */
    /* Builder::progress(ProgressBar) (core:70-75; feature `indicatif`): called on the building thread with the number of
     * points whose insertion has been enqueued so far (set_position, core:519-525) and the total (set_length, core:216-222);
     * the last call has done == total (finish, core:331-334).  NULL = no reporting. */
    void (*progress)(uint64_t done, uint64_t total, void* user);
    void*    progress_user;
} idb_params;

/* Builder::default() (core:101-113) — except `seed`, which the reference draws from entropy; here 0. */
IDB_API idb_status idb_params_default(idb_params* p);

/* Builder::build_hnsw(points) -> (Hnsw, Vec<PointId>) (core:83-85 -> Hnsw::new core:209-345).
 * rows: n x dim row-major host f32.  out_ids[i] = PointId assigned to input row i (core:262-270); may be NULL. */
IDB_API idb_status idb_build_f32(const float* rows, uint64_t n, uint32_t dim, const idb_params* params,
                         idb_index** out_index, uint32_t* out_ids);

/* "Search a given graph": adopt a graph built elsewhere (the reference, the oracle, a loaded .idx file).
 * This is the parity entry point.  Mirrors the fields of `Hnsw` (core:194-199):
 *   points   n x dim, PointId order                       (Hnsw::points)
 *   zero     n x 2M u32, INVALID-terminated rows          (Hnsw::zero, ZeroNode types:83-85)
 *   upper[l] upper_n[l] x M u32 for layer l+1             (Hnsw::layers, UpperNode types:63) */
IDB_API idb_status idb_index_from_graph_f32(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef_search,
                                    const uint32_t* zero, uint32_t n_upper, const uint32_t* const* upper,
                                    const uint64_t* upper_n, int32_t device, idb_index** out_index);

/* Same, but the point rows are rounded to bf16 and stored that way (BASELINE config 4's data format).  Results equal the f32
 * engine / the reference algorithm run on the bf16-rounded points. */
IDB_API idb_status idb_index_from_graph_bf16(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef_search,
                                     const uint32_t* zero, uint32_t n_upper, const uint32_t* const* upper,
                                     const uint64_t* upper_n, int32_t device, idb_index** out_index);

/* Hnsw::search(point, &mut Search) (core:352-383), batched: one independent search per query row.
 * The reference returns the whole `nearest` list (<= ef_search items, ascending by (distance, pid));
 * callers take the first k.  Here: out_ids/out_dist are nq x k (row q holds the first min(len,k) items,
 * padded with IDB_INVALID / +inf), out_len[q] = len(nearest) (what `ExactSizeIterator::len` reports).
 * ef_search == 0 uses the index's own ef_search (Hnsw::ef_search, core:195).  out_dist / out_len may be NULL.
 * Thread safety: `Hnsw<P>: Sync` (core:352-356) — any number of host threads may call this on one index at once; each call
 * takes an idle submission lane (own CUDA stream and control state, idb_index_num_lanes() of them) so concurrent callers overlap on
 * the device.  Buffers may be pageable or pinned (idb_host_alloc); results for pageable output buffers are staged through pinned
 * memory inside the library, so one caller's read-back never stalls another caller's launches.
 * Device-wide side effect: the per-warp visited tables of the traversal kernels live in a per-device pool shared by every index; the
 * first search/build on a device reserves part of the device's persisting-L2 set-aside for them (cudaLimitPersistingL2CacheSize, as
 * much as the tables in use need, at most the device maximum) and every launch carries an access-policy window for them as a launch
 * attribute (no stream state).  The reservation is returned when the last index on the device is freed.  A host application that
 * manages the persisting L2 itself calls idb_device_set_persisting_l2(device, 0) (or sets IDB_L2_PERSIST=0): results are identical,
 * throughput ~10 % lower. */
IDB_API idb_status idb_search_batch_f32(idb_index* index, const float* queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                uint32_t* out_ids, float* out_dist, uint32_t* out_len);

/* Same, with queries and outputs already resident in HBM (device pointers, same device as the index;
 * d_queries is nq x dim row-major).  Enqueues on idb_index_stream(index) (= lane 0) and returns without syncing.
 * A query that overflows an internal per-query structure even in the retry pass gets out_len = 0 and IDB_INVALID ids;
 * idb_last_search_failures() reports how many did (the host-buffer call returns IDB_ERR_CAPACITY instead). */
IDB_API idb_status idb_search_batch_device(idb_index* index, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                   uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len);
/* The same on submission lane `lane` < idb_index_num_lanes(): calls on one lane are stream-ordered (idb_index_lane_stream), calls
 * on different lanes overlap — the next batch's thread blocks move in as the previous batch's drain, so back-to-back batches
 * issued alternately on two lanes keep the GPU full across batch boundaries. */
IDB_API idb_status idb_search_batch_device_lane(idb_index* index, uint32_t lane, const float* d_queries, uint64_t nq, uint32_t ef_search,
                                        uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len);
IDB_API uint32_t idb_index_num_lanes(void);
IDB_API void* idb_index_lane_stream(idb_index* index, uint32_t lane);  /* cudaStream_t of that lane */
/* Waits for the last call on `lane` and reports how many of its queries failed even in the retry pass (0 = all results valid). */
IDB_API idb_status idb_last_search_failures(idb_index* index, uint32_t lane, uint32_t* out_failed);
/* Diagnostics: how many queries of the last call on `lane` overflowed their per-warp visited table / tie list in the main pass and were
 * re-run by the retry pass (their results are valid; a persistently non-zero figure costs throughput, and the library then switches the
 * index to its larger, DRAM-resident visited flavour by itself).  lane = 0xFFFFFFFF: the lane the last call on this index used. */
IDB_API idb_status idb_last_search_retried(idb_index* index, uint32_t lane, uint32_t* out_retried);
/* enabled = 0: this library never touches the device's persisting-L2 limit nor attaches access-policy windows on `device`. */
IDB_API idb_status idb_device_set_persisting_l2(int32_t device, int32_t enabled);

/* Per-query traversal counters of the LAST search call issued on this index (whichever lane it used; for the roofline accounting,
 * SURVEY §8d): out is nq x 4 u64 = {n_expand_upper, n_dist_upper, n_expand_zero, n_dist_zero}. */
IDB_API idb_status idb_last_search_counters(idb_index* index, uint64_t nq, uint64_t* out);

/* Introspection: Hnsw::iter / Index<PointId> (core:386-391, types:269-275) and the graph itself. */
typedef struct idb_info {
    uint64_t n;
    uint32_t dim;
    uint32_t M;
    uint32_t ef_search;
    uint32_t n_layers;          /* 0 for an empty index, else 1 + number of upper layers */
    uint64_t layer_n[32];       /* node count per layer, [0] = n */
    int32_t  device;
    uint32_t storage;           /* IDB_STORAGE_* */
} idb_info;
IDB_API idb_status idb_index_info(const idb_index* index, idb_info* out);
IDB_API idb_status idb_index_export_points(const idb_index* index, float* out /* n x dim */);
IDB_API idb_status idb_index_export_zero(const idb_index* index, uint32_t* out /* n x 2M */);
IDB_API idb_status idb_index_export_upper(const idb_index* index, uint32_t layer /* 1-based */, uint32_t* out /* n_l x M */);

/* Hnsw::dump / Hnsw::load of the Python binding (py:121-137): bincode-1.3 layout of `Hnsw{ef_search, points, zero, layers}`
 * (core:193-199).  dim and M are not stored in the file (fixed-size arrays in the reference: dim = 300, M = 32), so load takes
 * them.  *out_values_offset (may be NULL) = file offset where an HnswMap's `values` begin (core:131-134), or the file size. */
IDB_API idb_status idb_index_save(const idb_index* index, const char* path);
IDB_API idb_status idb_index_load(const char* path, uint32_t dim, uint32_t M, int32_t device, idb_index** out_index, uint64_t* out_values_offset);

/* Measurement hooks (bench.py): when enabled, CUDA events are recorded on the index stream immediately around the
 * dominant kernel of each call (K1 search_layer for searches); idb_index_last_kernel_ms waits for that kernel and
 * returns its duration and how many of this library's kernels the last call launched. */
IDB_API idb_status idb_index_set_profiling(idb_index* index, int32_t enabled);
IDB_API idb_status idb_index_last_kernel_ms(idb_index* index, float* out_ms, uint32_t* out_launches);

/* Measurement only: random point-row gathers in K1's launch shape and arithmetic (no visited set / adjacency / merge), to show
 * the gather ceiling of the device next to K1's own rate.  chain = independent 16-row batches between two dependent steps
 * (0 = all independent).  *out_bytes = bytes gathered per run. */
IDB_API idb_status idb_debug_gather_bench(idb_index* index, uint32_t n_items, uint32_t batches, uint32_t chain, uint32_t reps,
                                          float* out_ms, double* out_bytes);
/* Same with K1's second memory stream riding along: `atomics_per_batch` visited-style atomicAnd per 16-row batch on a per-warp
 * n-bit bitmap that is wiped after every item.  mode 1: atomics overlap the row loads (traffic-mix ceiling); mode 2: the row loads
 * wait for them (K1's dependency).  *out_bytes counts the ROW bytes only, like K1's algorithmic bytes. */
IDB_API idb_status idb_debug_gather_mix_bench(idb_index* index, uint32_t n_items, uint32_t batches, uint32_t chain, uint32_t reps,
                                              uint32_t atomics_per_batch, uint32_t mode, float* out_ms, double* out_bytes);

IDB_API void* idb_index_stream(idb_index* index);      /* lane 0's cudaStream_t: builds, uploads and idb_search_batch_device run on it */
IDB_API idb_status idb_index_sync(idb_index* index);   /* cudaStreamSynchronize on every lane of the index */
IDB_API void idb_index_free(idb_index* index);         /* Drop for Hnsw */

/* ---- Index sharded by PointId range across the GPUs of one box (one process per GPU) ------------------------------
 * The reference has no distributed path; this is north_star's layout: every rank owns an independent index over its
 * contiguous range of the input rows, every query is searched on every shard, and ONE ncclAllGather of the per-shard
 * top-k (packed (distance, global id) keys) is followed by a merge kernel.  Results: the k smallest (distance, global id)
 * of the union of the shards' `nearest` lists; identical on every rank. */
#define IDB_UNIQUE_ID_BYTES 128
typedef struct idb_comm idb_comm;
IDB_API idb_status idb_comm_unique_id(void* out_unique_id /* IDB_UNIQUE_ID_BYTES, made on one rank, shared by the host app */);
IDB_API idb_status idb_comm_create(const void* unique_id, int32_t rank, int32_t world, int32_t device, idb_comm** out);
IDB_API void idb_comm_free(idb_comm* comm);
/* global_ids[pid] = the caller's id of the row that became PointId pid on this shard (NULL clears the map). */
IDB_API idb_status idb_index_set_id_map(idb_index* index, const uint32_t* global_ids);
/* Collective over `comm`: every rank passes the same queries.  out_ids are GLOBAL ids. */
IDB_API idb_status idb_sharded_search_batch_f32(idb_index* shard, idb_comm* comm, const float* queries, uint64_t nq, uint32_t ef_search,
                                        uint32_t k, uint32_t* out_ids, float* out_dist, uint32_t* out_len);
IDB_API idb_status idb_sharded_search_batch_device(idb_index* shard, idb_comm* comm, const float* d_queries, uint64_t nq,
                                           uint32_t ef_search, uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len);
/* The same for a rank that holds SEVERAL shards on its device (e.g. 8 PointId ranges over 2 or 4 GPUs, or all 8 on one GPU: the
 * denominator of the 1 -> 8 GPU scaling figure).  The rank's shards are searched concurrently (one launch each, overlapping on the
 * device), their top-k lists pre-merged on the device, and the rank still contributes ONE k-list per query to the ONE all-gather. */
IDB_API idb_status idb_sharded_search_batch_f32_multi(idb_index* const* shards, uint32_t n_shards, idb_comm* comm, const float* queries,
                                              uint64_t nq, uint32_t ef_search, uint32_t k, uint32_t* out_ids, float* out_dist,
                                              uint32_t* out_len);
IDB_API idb_status idb_sharded_search_batch_device_multi(idb_index* const* shards, uint32_t n_shards, idb_comm* comm,
                                                 const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                                 uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len);

/* The canonical squared-L2 of one pair, evaluated on the device (used by parity tests; FloatArray::distance, py:378-421). */
IDB_API idb_status idb_distance_f32(const float* a, const float* b, uint32_t dim, int32_t device, float* out);

/* Pinned host memory helpers for callers that want true async H2D/D2H. */
IDB_API idb_status idb_host_alloc(size_t bytes, void** out);
IDB_API void idb_host_free(void* p);

IDB_API const char* idb_last_error(void);
IDB_API const char* idb_version(void);
IDB_API int32_t idb_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* INSTANT_DISTANCE_B200_H */
