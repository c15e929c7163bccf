// sharded.cu — the index sharded by PointId range across the GPUs of one box (BASELINE config 5).
//
// One process per GPU.  Every rank owns an independent HNSW over its contiguous range of the input rows (own shuffle, own
// PointId space, own entry point); every query is searched on every shard with the same ef; then ONE ncclAllGather of the
// per-shard top-k — packed as u64 keys (canonical distance bits << 32 | global row id) by K1's epilogue — and a merge
// kernel that keeps the k smallest keys of the union, ties broken by the lower global id.  No other collective anywhere.
// NCCL is bound at run time (dlopen of libnccl.so.2: the system 2.27 or whichever copy the host application already loaded).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.cuh"

namespace idb {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.handle, "ncclAllGather"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}

#define NCCL_TRY(expr)                                                                                                   \
    do {                                                                                                                 \
        ncclResult_t r__ = (expr);                                                                                       \
        if (r__ != ncclSuccess) return ::idb::fail(IDB_ERR_NCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl().GetErrorString(r__)); \
    } while (0)

// K4: merge of the G per-shard k-lists of one query (one warp per query): keep the k smallest keys of the union.
// Keys are unique (global ids are), so rank(key) = #{keys smaller} is a permutation.  The lists are NOT assumed sorted by
// the full key: a shard orders exact-distance ties by its local PointId, the merged order is by global id.
// out_keys != null: write the merged keys (the local pre-merge of a rank that holds several shards) instead of ids / distances.
__global__ void merge_topk_kernel(const uint64_t* all_keys /* G x nq x k */, uint32_t G, uint64_t nq, uint32_t k,
                                  uint32_t* out_ids, float* out_dist, uint32_t* out_len, uint64_t* out_keys) {
    extern __shared__ uint64_t sm_keys[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint64_t* keys = sm_keys + (size_t)warp * G * k;
    const uint32_t total = G * k;
    for (uint64_t q = (uint64_t)blockIdx.x * wpb + warp; q < nq; q += (uint64_t)gridDim.x * wpb) {
        for (uint32_t t = lane; t < total; t += 32) {
            const uint32_t g = t / k, j = t - g * k;
            keys[t] = all_keys[((size_t)g * nq + q) * k + j];
        }
        __syncwarp();
        uint32_t found = 0;
        for (uint32_t t = lane; t < total; t += 32) {
            const uint64_t key = keys[t];
            if (key == kKeyNone) continue;
            uint32_t rank = 0;
            for (uint32_t i = 0; i < total; ++i) rank += keys[i] < key ? 1u : 0u;
            if (rank < k) {
                if (out_keys) out_keys[q * k + rank] = key;
                else {
                    out_ids[q * k + rank] = (uint32_t)key;
                    if (out_dist) out_dist[q * k + rank] = __uint_as_float((uint32_t)(key >> 32));
                }
            }
        }
        // number of real results = min(k, total non-empty keys); pad the tail
        uint32_t real = 0;
        for (uint32_t t = lane; t < total; t += 32) real += keys[t] != kKeyNone ? 1u : 0u;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) real += __shfl_xor_sync(kFullMask, real, off);
        found = min(real, k);
        for (uint32_t j = found + lane; j < k; j += 32) {
            if (out_keys) out_keys[q * k + j] = kKeyNone;
            else {
                out_ids[q * k + j] = kInvalid;
                if (out_dist) out_dist[q * k + j] = __int_as_float(0x7f800000);
            }
        }
        if (out_len && lane == 0) out_len[q] = found;
        __syncwarp();
    }
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

}  // namespace idb

using namespace idb;

extern "C" {

idb_status idb_comm_unique_id(void* out) {
    if (!out) return fail(IDB_ERR_INVALID_ARG, "out is null");
    if (!nccl().ok) return fail(IDB_ERR_NCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    static_assert(IDB_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
    return IDB_OK;
}

idb_status idb_comm_create(const void* unique_id, int32_t rank, int32_t world, int32_t device, idb_comm** out) {
    if (!unique_id || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(IDB_ERR_INVALID_ARG, "rank %d / world %d invalid", rank, world);
    if (!nccl().ok) return fail(IDB_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CUDA_TRY(cudaSetDevice(device));
    auto* c = new (std::nothrow) Comm();
    if (!c) return fail(IDB_ERR_OOM, "host allocation failed");
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(IDB_ERR_NCCL, "ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    }
    *out = reinterpret_cast<idb_comm*>(c);
    return IDB_OK;
}

void idb_comm_free(idb_comm* comm) {
    Comm* c = reinterpret_cast<Comm*>(comm);
    if (!c) return;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

idb_status idb_index_set_id_map(idb_index* index, const uint32_t* global_ids) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    if (!global_ids) {
        cudaFree(ix->d_id_map);
        ix->d_id_map = nullptr;
        return IDB_OK;
    }
    if (ix->n == 0) return IDB_OK;
    if (!ix->d_id_map) CUDA_TRY(cudaMalloc(&ix->d_id_map, ix->n * 4));
    CUDA_TRY(cudaMemcpyAsync(ix->d_id_map, global_ids, ix->n * 4, cudaMemcpyHostToDevice, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return IDB_OK;
}

}  // extern "C" (reopened below)

namespace idb {
idb_status search_device_keys(Index* ix, Lane& ln, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k, uint32_t* d_ids,
                              uint64_t* d_keys);  // api.cu

static idb_status launch_merge(Index* ix, cudaStream_t st, const uint64_t* keys, uint32_t G, uint64_t nq, uint32_t k, uint32_t* d_ids,
                               float* d_dist, uint32_t* d_len, uint64_t* d_keys, int max_smem) {
    const size_t per_warp = (size_t)G * k * 8;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > (size_t)max_smem) wpb >>= 1;
    const size_t smem = (size_t)wpb * per_warp;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<uint64_t>((nq + wpb - 1) / wpb, (uint64_t)ix->num_sms * 8);
    merge_topk_kernel<<<grid, wpb * 32, smem, st>>>(keys, G, nq, k, d_ids, d_dist, d_len, d_keys);
    CUDA_TRY(cudaGetLastError());
    return IDB_OK;
}

// The rank's shards each run K1 on their own lane-0 stream (K1's epilogue packs (distance bits, global id) keys; the launches overlap
// on the device: one table pool, the next shard's thread blocks move in as the previous shard's drain) -> local pre-merge when the
// rank holds more than one shard -> ONE ncclAllGather -> merge kernel.  Main stream = lane 0 of the first shard.
// The caller holds lanes[0].mu of every shard.
static idb_status sharded_search_locked(Index* const* shards, uint32_t n_local, Comm* c, const float* d_queries, uint64_t nq,
                                        uint32_t ef_search, uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    Index* ix = shards[0];
    Lane& ln = ix->lanes[0];
    CUDA_TRY(cudaSetDevice(ix->device));
    // merge kernels: world * k (and n_local * k) keys per query in shared memory; check the launches BEFORE anything is enqueued
    int max_smem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ix->device));
    const uint64_t widest = std::max<uint64_t>((uint64_t)c->world, n_local) * k * 8;
    if (widest > (uint64_t)max_smem)
        return fail(IDB_ERR_UNSUPPORTED, "%llu keys per query do not fit the merge kernel's shared memory (%d bytes)",
                    (unsigned long long)(widest / 8), max_smem);
    const size_t per = (size_t)nq * k;
    CUDA_TRY(ensure_u64(ln.keys_local, ln.keys_local_cap, per * (n_local > 1 ? n_local + 1 : 1)));
    CUDA_TRY(ensure_u64(ln.keys_all, ln.keys_all_cap, per * c->world));
    uint64_t* shard_keys = n_local > 1 ? ln.keys_local + per : ln.keys_local;  // [n_local][nq][k]; the pre-merge writes keys_local[0..per)
    cudaEvent_t fork = nullptr;
    if (n_local > 1) {
        CUDA_TRY(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(fork, ln.stream));  // the queries (and anything else the caller enqueued) are ready
    }
    idb_status st = IDB_OK;
    for (uint32_t i = 0; i < n_local && st == IDB_OK; ++i) {
        Index* sx = shards[i];
        Lane& sl = sx->lanes[0];
        if (i > 0 && cudaStreamWaitEvent(sl.stream, fork, 0) != cudaSuccess) st = fail(IDB_ERR_CUDA, "cudaStreamWaitEvent failed");
        if (st == IDB_OK && cudaSuccess != ensure_u32(sl.ids, sl.ids_cap, per)) st = fail(IDB_ERR_OOM, "scratch allocation failed");
        if (st == IDB_OK) st = search_device_keys(sx, sl, d_queries, nq, ef_search, k, sl.ids, shard_keys + (size_t)i * per);
        if (st == IDB_OK && i > 0) {  // join: the main stream continues after this shard's K1
            cudaEvent_t done = nullptr;
            if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(done, sl.stream) != cudaSuccess ||
                cudaStreamWaitEvent(ln.stream, done, 0) != cudaSuccess)
                st = fail(IDB_ERR_CUDA, "stream join failed");
            if (done) cudaEventDestroy(done);  // (released once the recorded work has completed)
        }
    }
    if (fork) cudaEventDestroy(fork);
    if (st != IDB_OK) return st;
    if (n_local > 1) {
        st = launch_merge(ix, ln.stream, shard_keys, n_local, nq, k, nullptr, nullptr, nullptr, ln.keys_local, max_smem);
        if (st != IDB_OK) return st;
    }
    NCCL_TRY(nccl().AllGather(ln.keys_local, ln.keys_all, per, ncclUint64, c->comm, ln.stream));
    st = launch_merge(ix, ln.stream, ln.keys_all, (uint32_t)c->world, nq, k, d_out_ids, d_out_dist, d_out_len, nullptr, max_smem);
    if (st != IDB_OK) return st;
    ln.last_launches = 2 * n_local + (n_local > 1 ? 1 : 0) + 2;  // K1 + retry per shard, pre-merge, all-gather, merge
    return IDB_OK;
}

// Validates the shard list and locks lane 0 of every shard (in address order: two callers with the same shards cannot deadlock).
struct ShardLocks {
    std::vector<Index*> order;
    ~ShardLocks() { for (auto it = order.rbegin(); it != order.rend(); ++it) (*it)->lanes[0].mu.unlock(); }
    idb_status lock(idb_index* const* shards, uint32_t n) {
        if (!shards || n == 0 || n > 64) return fail(IDB_ERR_INVALID_ARG, "shards: need 1..64 index handles");
        std::vector<Index*> v;
        for (uint32_t i = 0; i < n; ++i) {
            Index* ix = reinterpret_cast<Index*>(shards[i]);
            if (!ix) return fail(IDB_ERR_INVALID_ARG, "shard %u is null", i);
            if (ix->device != reinterpret_cast<Index*>(shards[0])->device || ix->dim != reinterpret_cast<Index*>(shards[0])->dim)
                return fail(IDB_ERR_INVALID_ARG, "shard %u: all shards of a rank must live on one device and have one dim", i);
            v.push_back(ix);
        }
        std::sort(v.begin(), v.end());
        if (std::adjacent_find(v.begin(), v.end()) != v.end()) return fail(IDB_ERR_INVALID_ARG, "the same index is listed twice");
        for (Index* ix : v) { ix->lanes[0].mu.lock(); order.push_back(ix); ix->last_lane.store(0); }
        return IDB_OK;
    }
};
}  // namespace idb

extern "C" {

// All device pointers; d_queries is nq x dim.  Collective: every rank of `comm` calls it with the same queries.
idb_status idb_sharded_search_batch_device_multi(idb_index* const* shards, uint32_t n_shards, idb_comm* comm, const float* d_queries,
                                                 uint64_t nq, uint32_t ef_search, uint32_t k, uint32_t* d_out_ids, float* d_out_dist,
                                                 uint32_t* d_out_len) {
    if (!comm) return fail(IDB_ERR_INVALID_ARG, "comm is null");
    if (nq == 0) return IDB_OK;
    if (!d_queries || !d_out_ids || k == 0) return fail(IDB_ERR_INVALID_ARG, "bad argument");
    ShardLocks locks;
    idb_status st = locks.lock(shards, n_shards);
    if (st != IDB_OK) return st;
    return sharded_search_locked(reinterpret_cast<Index* const*>(shards), n_shards, reinterpret_cast<Comm*>(comm), d_queries, nq, ef_search,
                                 k, d_out_ids, d_out_dist, d_out_len);
}

idb_status idb_sharded_search_batch_device(idb_index* index, idb_comm* comm, const float* d_queries, uint64_t nq, uint32_t ef_search,
                                           uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    return idb_sharded_search_batch_device_multi(&index, 1, comm, d_queries, nq, ef_search, k, d_out_ids, d_out_dist, d_out_len);
}

idb_status idb_sharded_search_batch_f32_multi(idb_index* const* shards, uint32_t n_shards, idb_comm* comm, const float* queries, uint64_t nq,
                                              uint32_t ef_search, uint32_t k, uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    if (!comm) return fail(IDB_ERR_INVALID_ARG, "comm is null");
    if (nq == 0) return IDB_OK;
    if (!queries || !out_ids || k == 0) return fail(IDB_ERR_INVALID_ARG, "bad argument");
    ShardLocks locks;  // held across staging, search and copy-back: the staging buffers belong to this call
    idb_status st = locks.lock(shards, n_shards);
    if (st != IDB_OK) return st;
    Index* ix = reinterpret_cast<Index*>(shards[0]);
    Lane& ln = ix->lanes[0];
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(ensure_f32(ln.q2, ln.q2_cap, nq * ix->dim));
    CUDA_TRY(ensure_u32(ln.ids2, ln.ids2_cap, nq * k));
    CUDA_TRY(ensure_f32(ln.dist, ln.dist_cap, nq * k));
    CUDA_TRY(ensure_u32(ln.len, ln.len_cap, nq));
    CUDA_TRY(cudaMemcpyAsync(ln.q2, queries, nq * ix->dim * 4, cudaMemcpyHostToDevice, ln.stream));
    st = sharded_search_locked(reinterpret_cast<Index* const*>(shards), n_shards, reinterpret_cast<Comm*>(comm), ln.q2, nq, ef_search, k,
                               ln.ids2, ln.dist, ln.len);
    if (st != IDB_OK) return st;
    HostOut ho;  // (pageable output buffers are staged through pinned memory: internal.cuh)
    ho.add(out_ids, ln.ids2, nq * k * 4);
    ho.add(out_dist, ln.dist, nq * k * 4);
    ho.add(out_len, ln.len, nq * 4);
    CUDA_TRY(ho.enqueue(ln));
    CUDA_TRY(cudaStreamSynchronize(ln.stream));  // every shard's stream was joined into this one
    ho.finish();
    uint32_t failed = 0;
    for (uint32_t i = 0; i < n_shards; ++i) {
        Lane& sl = reinterpret_cast<Index*>(shards[i])->lanes[0];
        uint32_t ctrl[16] = {0};
        if (sl.ctrl && sl.last_nq) CUDA_TRY(cudaMemcpy(ctrl, sl.ctrl, 64, cudaMemcpyDeviceToHost));
        failed += ctrl[12];
    }
    if (failed) return fail(IDB_ERR_CAPACITY, "%u queries overflowed an internal per-query structure on this rank's shards", failed);
    return IDB_OK;
}

idb_status idb_sharded_search_batch_f32(idb_index* index, idb_comm* comm, const float* queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                        uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    return idb_sharded_search_batch_f32_multi(&index, 1, comm, queries, nq, ef_search, k, out_ids, out_dist, out_len);
}

}  // extern "C"
