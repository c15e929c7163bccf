// build.cu — host driver of Builder::build_hnsw (instant-distance/src/lib.rs:83-85 -> Hnsw::new, lib.rs:209-345) on the GPU.
//
//   1. layer schedule            lib.rs:238-250  (f32 multiply, truncating cast)
//   2. seeded shuffle            lib.rs:257-270  (xoshiro256++ seeded through SplitMix64; widening-multiply range sampling —
//                                                 the rand crate is not vendored in the reference tree: parity unpinned)
//   3. per layer, top first      lib.rs:304-329  batches of concurrent inserts (KA -> K2 -> sort -> K2'), then the
//                                                 UpperNode snapshot (K5)
// The batch schedule replaces rayon: batch = min(16384, max(1, inserted / 8)); the top layer is sequential like the
// reference's (lib.rs:313-314).  insert_batch = 1 reproduces the sequential reference order exactly.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "build_dispatch.cuh"

namespace idb {

cudaError_t build_dispatch_ch1(const BuildArgs&, const BuildLaunch&, cudaStream_t);
cudaError_t build_dispatch_ch2(const BuildArgs&, const BuildLaunch&, cudaStream_t);
cudaError_t build_dispatch_ch3(const BuildArgs&, const BuildLaunch&, cudaStream_t);
cudaError_t build_dispatch_ch4(const BuildArgs&, const BuildLaunch&, cudaStream_t);
cudaError_t build_dispatch_ch6(const BuildArgs&, const BuildLaunch&, cudaStream_t);
cudaError_t build_dispatch_ch8(const BuildArgs&, const BuildLaunch&, cudaStream_t);
cudaError_t build_dispatch_long(const BuildArgs&, const BuildLaunch&, cudaStream_t);

static cudaError_t build_dispatch_any(int ch, const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) {
    switch (ch) {
        case 1: return build_dispatch_ch1(a, l, st);
        case 2: return build_dispatch_ch2(a, l, st);
        case 3: return build_dispatch_ch3(a, l, st);
        case 4: return build_dispatch_ch4(a, l, st);
        case 5: case 6: return build_dispatch_ch6(a, l, st);
        case 7: case 8: return build_dispatch_ch8(a, l, st);
        default: return build_dispatch_long(a, l, st);
    }
}

namespace {

// No heuristic (Builder::select_heuristic(None), lib.rs:466-469): found = first min(len, 2M) of `nearest`.
__global__ void select_simple_kernel(BuildArgs a) {
    const uint32_t cap = 2 * a.g.M;
    for (uint32_t w = blockIdx.x; w < a.count; w += gridDim.x) {
        const uint32_t neu = a.base + w;
        const uint32_t total = min(a.cand_cnt[w], cap);
        uint32_t* row = a.zero + (size_t)neu * cap;
        for (uint32_t t = threadIdx.x; t < cap; t += blockDim.x) {
            const uint32_t pid = t < total ? key_pid(a.cand_keys[(size_t)w * a.cand_cap + t]) : kInvalid;
            row[t] = pid;
            a.pairs[(size_t)w * cap + t] = t < total ? (((uint64_t)pid << 32) | neu) : kKeyNone;
        }
    }
}

// Segment heads of the sorted link requests: one work item per distinct target row.
__global__ void segment_heads_kernel(const uint64_t* sorted_pairs, uint32_t n, uint32_t* seg_start, uint32_t* n_seg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t tgt = (uint32_t)(sorted_pairs[i] >> 32);
    if (tgt == kInvalid) return;
    if (i == 0 || (uint32_t)(sorted_pairs[i - 1] >> 32) != tgt) seg_start[atomicAdd(n_seg, 1u)] = i;
}

// K5: UpperNode::from_zero (types.rs:65-71) for nodes [0, n_l).
__global__ void snapshot_kernel(const uint32_t* zero, uint32_t* upper, uint64_t n_l, uint32_t M) {
    const uint64_t total = n_l * M;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = i / M, e = i - v * M;
        upper[i] = zero[v * 2 * M + e];
    }
}

// K6: points[rank] = rows[order[rank]]  (lib.rs:263-270), zero-padding each row to a multiple of 4 floats.
__global__ void gather_rows_kernel(const float* rows, const uint32_t* order, float* points, uint64_t n, uint32_t dim, uint32_t stride) {
    const uint64_t total = n * stride;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / stride;
        const uint32_t c = (uint32_t)(i - r * stride);
        points[i] = c < dim ? rows[(uint64_t)order[r] * dim + c] : 0.f;
    }
}

struct Xoshiro256pp {  // rand's SmallRng on 64-bit targets
    uint64_t s[4];
    explicit Xoshiro256pp(uint64_t seed) {  // SeedableRng::seed_from_u64: SplitMix64 stream
        uint64_t state = seed;
        for (int i = 0; i < 4; ++i) {
            state += 0x9e3779b97f4a7c15ull;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        const uint64_t res = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return res;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    uint32_t below(uint32_t range) {  // random_range(0..range)
        const uint64_t m = (uint64_t)next_u32() * range;
        uint32_t result = (uint32_t)(m >> 32);
        const uint32_t lo = (uint32_t)m;
        if (lo > (uint32_t)(0u - range)) {
            const uint32_t hi2 = (uint32_t)(((uint64_t)next_u32() * range) >> 32);
            result += (uint32_t)(lo + hi2 < lo);
        }
        return result;
    }
};

// (size, cumulative) per layer, top layer first (lib.rs:238-249)
std::vector<std::pair<uint64_t, uint64_t>> layer_sizes(uint64_t n, uint32_t M, float ml) {
    std::vector<std::pair<uint64_t, uint64_t>> sizes;
    uint64_t num = n;
    for (;;) {
        const float f = (float)num * ml;
        const uint64_t next = f >= 1.8446744e19f ? UINT64_MAX : (f > 0.0f ? (uint64_t)f : 0);
        if (next < M || next >= num) break;
        sizes.push_back({num - next, num});
        num = next;
    }
    sizes.push_back({num, num});
    std::reverse(sizes.begin(), sizes.end());
    return sizes;
}

struct BuildScratch {
    uint64_t* cand_keys = nullptr;
    uint32_t* cand_cnt = nullptr;
    uint64_t* pairs = nullptr;
    uint64_t* sorted = nullptr;
    uint32_t* seg_start = nullptr;
    uint32_t* status = nullptr;
    uint32_t* fail_list = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    // [0..8) KA work counter, [8..16) relink work counter, [16..20) n_seg, [20..24) KA fail count, [24..32) KA-retry work counter
    // (all reset per batch); [32..36) inserts that failed even in the retry pass (never reset)
    unsigned char* ctrl = nullptr;
    uint32_t* h_fail = nullptr;     // pinned, 2 words: [0] failed even in the retry pass, [1] KA overflows of the last batch
    ~BuildScratch() {
        cudaFree(cand_keys); cudaFree(cand_cnt); cudaFree(pairs); cudaFree(sorted); cudaFree(seg_start); cudaFree(status);
        cudaFree(fail_list); cudaFree(cub_tmp); cudaFree(ctrl);
        if (h_fail) cudaFreeHost(h_fail);
    }
};

}  // namespace

idb_status build_index(Index* ix, const float* rows, uint64_t n, uint32_t dim, const idb_params& p, uint32_t* out_ids) {
    const uint32_t M = p.M;
    ix->n = n;
    ix->dim = dim;
    ix->M = M;
    ix->ef_search = p.ef_search;
    ix->nchunks = (dim + 3) / 4;
    if (n == 0) return IDB_OK;  // lib.rs:224-234
    cudaStream_t st = ix->stream;
    const uint32_t cap = 2 * M;
    const size_t stride = (size_t)ix->nchunks * 4;

    // ---- 1. layers, 2. shuffle ------------------------------------------------------------------------------
    const auto sizes = layer_sizes(n, M, p.ml);
    const uint32_t num_layers = (uint32_t)sizes.size(), top = num_layers - 1;
    if (top > 31) return fail(IDB_ERR_INVALID_ARG, "ml = %g produces %u layers (max 32)", (double)p.ml, num_layers);
    std::vector<uint32_t> order(n);
    {
        Xoshiro256pp rng(p.seed);
        std::vector<std::pair<uint32_t, uint64_t>> sh(n);
        for (uint64_t i = 0; i < n; ++i) sh[i] = {rng.below((uint32_t)n), i};
        std::sort(sh.begin(), sh.end());
        for (uint64_t r = 0; r < n; ++r) {
            order[r] = (uint32_t)sh[r].second;
            if (out_ids) out_ids[sh[r].second] = (uint32_t)r;
        }
    }

    // ---- device arrays ----------------------------------------------------------------------------------------
    {
        float* d_rows = nullptr;
        uint32_t* d_order = nullptr;
        CUDA_TRY(cudaMalloc(&ix->d_points, n * stride * sizeof(float)));
        CUDA_TRY(cudaMalloc(&d_rows, n * (size_t)dim * sizeof(float)));
        CUDA_TRY(cudaMalloc(&d_order, n * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpyAsync(d_rows, rows, n * (size_t)dim * sizeof(float), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_order, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        gather_rows_kernel<<<ix->num_sms * 8, 256, 0, st>>>(d_rows, d_order, ix->d_points, n, dim, (uint32_t)stride);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(d_rows);
        cudaFree(d_order);
    }
    if (p.storage == IDB_STORAGE_BF16) {
        idb_status sb = ix->narrow_points_to_bf16();
        if (sb != IDB_OK) return sb;
    }
    CUDA_TRY(cudaMalloc(&ix->d_zero, n * (size_t)cap * 4));
    CUDA_TRY(fill_u32(ix->d_zero, n * (size_t)cap, kInvalid, st));
    std::vector<const uint32_t*> ptrs;
    for (uint32_t l = 1; l <= top; ++l) {
        const uint64_t n_l = sizes[num_layers - 1 - l].second;
        uint32_t* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, std::max<size_t>(4, n_l * (size_t)M * 4)));
        ix->d_upper.push_back(d);
        ix->upper_n.push_back(n_l);
        ptrs.push_back(d);
    }
    CUDA_TRY(cudaMalloc(&ix->d_upper_ptrs, std::max<size_t>(1, top) * sizeof(uint32_t*)));
    if (top) CUDA_TRY(cudaMemcpyAsync(ix->d_upper_ptrs, ptrs.data(), top * sizeof(uint32_t*), cudaMemcpyHostToDevice, st));

    // ---- batch schedule + scratch -------------------------------------------------------------------------------
    const uint32_t efc = p.ef_construction;
    // Defaults from the sweep in profiles/r01_call11/12_tune_build_1M.jsonl (1M x 128): batch <= 16384 and <= 1/8 of the graph
    // keeps recall@10 within 0.002 of the reference algorithm's own graph (0.9713 vs 0.9729) at 2.5x the speed of batch 4096.
    uint32_t max_batch = p.insert_batch ? p.insert_batch : 16384u;
    uint32_t growth = 8;  // a batch never exceeds 1/8 of the graph it is inserted into
    if (!p.insert_batch) {  // tuning knobs for experiments (results stay valid HNSW graphs; determinism per setting)
        if (const char* e = std::getenv("IDB_BUILD_MAXBATCH")) max_batch = (uint32_t)std::max(1, std::atoi(e));
        if (const char* e = std::getenv("IDB_BUILD_GROWTH")) growth = (uint32_t)std::max(1, std::atoi(e));
    }
    const uint32_t cand_cap = std::max<uint32_t>((efc + 31) / 32 * 32, cap + kNewCap);
    BuildScratch bs;
    CUDA_TRY(cudaMalloc(&bs.cand_keys, (size_t)max_batch * cand_cap * 8));
    CUDA_TRY(cudaMalloc(&bs.cand_cnt, (size_t)max_batch * 4));
    CUDA_TRY(cudaMalloc(&bs.pairs, (size_t)max_batch * cap * 8));
    CUDA_TRY(cudaMalloc(&bs.sorted, (size_t)max_batch * cap * 8));
    CUDA_TRY(cudaMalloc(&bs.seg_start, (size_t)max_batch * cap * 4));
    CUDA_TRY(cudaMalloc(&bs.status, (size_t)max_batch * 4));
    CUDA_TRY(cudaMalloc(&bs.fail_list, (size_t)max_batch * 4));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&bs.h_fail), 8, cudaHostAllocDefault));
    CUDA_TRY(cudaMalloc(&bs.ctrl, 64));
    CUDA_TRY(cudaMemsetAsync(bs.ctrl, 0, 64, st));
    CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, bs.cub_bytes, bs.pairs, bs.sorted, (int)(max_batch * cap), 0, 64, st));
    CUDA_TRY(cudaMalloc(&bs.cub_tmp, std::max<size_t>(bs.cub_bytes, 16)));

    const int ch = (int)((ix->nchunks + 31) / 32);
    // Staging the kept rows in shared memory (72 KB per 2-warp CTA -> 6 warps per SM) measured 1.4x SLOWER than reading them
    // through L1/L2 with 32 resident warps (profiles/r01_call11_tune_build_1M.jsonl), so it is off unless asked for.
    bool stage = false;
    if (const char* e = std::getenv("IDB_BUILD_STAGE")) stage = std::atoi(e) != 0 && SelectSmem::bytes(cand_cap, M, ix->nchunks, true) <= 56 * 1024;
    const uint32_t k2_smem = (uint32_t)SelectSmem::bytes(cand_cap, M, ix->nchunks, stage);
    // K2 / K2' are persistent grids of 2-warp CTAs.  ncu (profiles/r02_call10_k2p_ncu_raw.csv): 64 registers, issue-bound (68 % issue active
    // with 16 warps per SM), so the grid asks for as many CTAs per SM as shared memory and registers allow, up to 16 (32 warps per SM).
    int k2_ctas_per_sm = (int)std::min<uint64_t>(16, std::max<uint64_t>(1, (200 * 1024) / std::max<uint32_t>(1, k2_smem * kBuildWarps)));
    if (const char* e = std::getenv("IDB_BUILD_CTAS")) k2_ctas_per_sm = std::max(1, std::atoi(e));

    BuildArgs a;
    std::memset(&a, 0, sizeof(a));
    a.g = ix->view();
    a.g.n_upper = top;
    a.zero = ix->d_zero;
    a.efc = efc;
    a.cand_cap = cand_cap;
    a.keep_pruned = p.keep_pruned ? 1u : 0u;
    a.cand_keys = bs.cand_keys;
    a.cand_cnt = bs.cand_cnt;
    a.pairs = bs.pairs;
    a.status = bs.status;
    a.fail_count = reinterpret_cast<uint32_t*>(bs.ctrl + 20);
    a.fail_list = bs.fail_list;
    a.sorted_pairs = bs.sorted;
    a.seg_start = bs.seg_start;
    a.n_seg = reinterpret_cast<uint32_t*>(bs.ctrl + 16);

    for (uint32_t li = 0; li < num_layers; ++li) {  // lib.rs:304-329
        const uint32_t layer = num_layers - li - 1;
        const uint64_t size = sizes[li].first, cumulative = sizes[li].second;
        const uint64_t start = std::max<uint64_t>(cumulative - size, 1), end = cumulative;
        uint64_t g0 = start;
        while (g0 < end) {
            uint64_t b = 1;
            if (layer != top && max_batch > 1) b = std::min<uint64_t>(max_batch, std::max<uint64_t>(1, g0 / growth));
            b = std::min<uint64_t>(b, end - g0);
            a.base = (uint32_t)g0;
            a.count = (uint32_t)b;
            a.layer = layer;
            a.n_pairs_cap = (uint32_t)(b * cap);
            CUDA_TRY(cudaMemsetAsync(bs.ctrl, 0, 32, st));
            // KA: descent of every insert, then (device-side, normally a no-op) a retry pass with 2^18-slot hash sets and 64k-entry
            // tie lists for the inserts whose per-warp structures overflowed (e.g. inside a cluster of thousands of duplicate vectors)
            BuildLaunch l;
            l.op = kOpInsertSearch;
            l.row_t = (int)((cap + 31) / 32);
            l.ef_t = (int)((efc + 31) / 32);
            l.stage = stage;
            l.smem_per_warp = k2_smem;
            l.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((b + kSearchWarps - 1) / kSearchWarps, (uint64_t)ix->search_grid()));
            {
                std::lock_guard<std::mutex> lk(ix->ctx->mu);  // the pool's tables must not be regrown under these launches
                SearchArgs tier;
                idb_status ts = ix->select_visited_tier(efc, tier, l.win);
                if (ts == IDB_OK) ts = ix->attach_window(ix->lanes[0], l.win);
                if (ts != IDB_OK) return ts;
                a.pool = tier.pool;
                a.gslots = tier.gslots;
                a.gshift = tier.gshift;
                a.vis_mode = tier.vis_mode;
                a.b16_cap_ids = tier.b16_cap_ids;
                a.b16_nb = tier.b16_nb;
                a.work_counter = reinterpret_cast<unsigned long long*>(bs.ctrl);
                CUDA_TRY(build_dispatch_any(ch, a, l, st));
                BuildArgs r = a;
                BuildLaunch lr = l;
                r.work_list = bs.fail_list;
                r.n_work_dev = a.fail_count;
                r.fail_list = nullptr;
                r.fail_count = reinterpret_cast<uint32_t*>(bs.ctrl + 32);
                r.work_counter = reinterpret_cast<unsigned long long*>(bs.ctrl + 24);
                r.pool = ix->ctx->retry_pool();
                r.gslots = kRetrySlots;
                r.gshift = 32 - 18;
                r.vis_mode = kVisHash;
                lr.grid = kRetryCtas;
                lr.win = LaunchWindow();
                CUDA_TRY(build_dispatch_any(ch, r, lr, st));
            }
            CUDA_TRY(cudaMemcpyAsync(bs.h_fail, bs.ctrl + 32, 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(bs.h_fail + 1, bs.ctrl + 20, 4, cudaMemcpyDeviceToHost, st));
            const int ka_b16 = a.vis_mode == kVisB16 ? ix->b16_level : 0;
            // K2: neighbour selection for the new nodes, own rows, link requests
            if (p.heuristic) {
                l.op = kOpSelectNew;
                l.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((b + kBuildWarps - 1) / kBuildWarps, (uint64_t)ix->num_sms * k2_ctas_per_sm));
                CUDA_TRY(build_dispatch_any(ch, a, l, st));
            } else {
                select_simple_kernel<<<(unsigned)std::min<uint64_t>(b, 1024), 64, 0, st>>>(a);
                CUDA_TRY(cudaGetLastError());
            }
            // group the link requests by target row
            size_t tmp = bs.cub_bytes;
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(bs.cub_tmp, tmp, bs.pairs, bs.sorted, (int)(b * cap), 0, 64, st));
            segment_heads_kernel<<<(unsigned)((b * cap + 255) / 256), 256, 0, st>>>(bs.sorted, (uint32_t)(b * cap), bs.seg_start,
                                                                                  reinterpret_cast<uint32_t*>(bs.ctrl + 16));
            CUDA_TRY(cudaGetLastError());
            // K2': re-prune every target row once
            l.op = p.heuristic ? kOpRelink : kOpRelinkSimple;
            l.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((b * cap + kBuildWarps - 1) / kBuildWarps, (uint64_t)ix->num_sms * k2_ctas_per_sm));
            a.work_counter = reinterpret_cast<unsigned long long*>(bs.ctrl + 8);
            CUDA_TRY(build_dispatch_any(ch, a, l, st));
            g0 += b;
            CUDA_TRY(cudaStreamSynchronize(st));  // fail fast: an insert that overflowed even the retry pass ends the build here
            if (*bs.h_fail)
                return fail(IDB_ERR_CAPACITY, "%u inserts overflowed an internal per-insert structure (visited table / tie list) in the batch ending at %llu",
                            *bs.h_fail, (unsigned long long)g0);
            ix->note_overflows(efc, b, bs.h_fail[1], ka_b16);  // too many b16 overflows: later batches use a larger flavour
            if (p.progress) p.progress(g0, n, p.progress_user);  // set_position (core:519-525)
        }
        if (layer != 0) {  // lib.rs:323-328
            snapshot_kernel<<<ix->num_sms * 4, 256, 0, st>>>(ix->d_zero, ix->d_upper[layer - 1], end, M);
            CUDA_TRY(cudaGetLastError());
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (p.progress) p.progress(n, n, p.progress_user);  // finish (core:331-334)
    return IDB_OK;
}

}  // namespace idb

using namespace idb;

extern "C" idb_status idb_build_f32(const float* rows, uint64_t n, uint32_t dim, const idb_params* params, idb_index** out_index,
                                    uint32_t* out_ids) {
    if (!out_index) return fail(IDB_ERR_INVALID_ARG, "out_index is null");
    *out_index = nullptr;
    if (!params) return fail(IDB_ERR_INVALID_ARG, "params is null");
    if (dim == 0) return fail(IDB_ERR_INVALID_ARG, "dim must be >= 1");
    if (n && !rows) return fail(IDB_ERR_INVALID_ARG, "rows is null");
    if (params->M < 2 || params->M > 64) return fail(IDB_ERR_INVALID_ARG, "M = %u unsupported (2..64)", params->M);
    if (n >= 0xFFFFFFFFull) return fail(IDB_ERR_INVALID_ARG, "N = %llu >= u32::MAX (lib.rs:256)", (unsigned long long)n);
    if (params->ef_construction == 0 || params->ef_construction > 1024)
        return fail(IDB_ERR_UNSUPPORTED, "ef_construction = %u unsupported (1..1024)", params->ef_construction);
    if (dim > 10240) return fail(IDB_ERR_UNSUPPORTED, "dim %u > 10240 is not supported (the owner row of a long-row traversal lives in shared memory)", dim);
    if (!(params->ml > 0.0f) || params->ml >= 1.0f) return fail(IDB_ERR_INVALID_ARG, "ml must be in (0, 1)");
    if (params->storage != IDB_STORAGE_F32 && params->storage != IDB_STORAGE_BF16) return fail(IDB_ERR_INVALID_ARG, "unknown storage %u", params->storage);
    if (params->heuristic && params->extend_candidates)
        return fail(IDB_ERR_UNSUPPORTED,
                    "Heuristic::extend_candidates = true is not supported: in the reference it re-locks the row being inserted "
                    "(lib.rs:438 write lock vs lib.rs:649 read lock through types.rs:146) and never returns");
    auto* ix = new (std::nothrow) Index();
    if (!ix) return fail(IDB_ERR_OOM, "host allocation failed");
    idb_status st = ix->init_device(params->device);
    if (st == IDB_OK) {
        std::lock_guard<std::mutex> lk(ix->mu);
        st = build_index(ix, rows, n, dim, *params, out_ids);
    }
    if (st != IDB_OK) { delete ix; return st; }
    *out_index = reinterpret_cast<idb_index*>(ix);
    return IDB_OK;
}
