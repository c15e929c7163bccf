// idb_build_graph — builds an HNSW graph on the GPU through the public C ABI and writes it to raw files.
//
// A plain C++ consumer of include/instant_distance_b200.h (no CUDA, no Python): the setup step bench.py's reference arm uses to
// obtain the graph both arms search (Builder::build is untimed setup there; the CPU arm itself never loads the CUDA library),
// and a minimal example of driving the library from a host language over the C ABI.
//
//   idb_build_graph <points.f32> <n> <dim> <M> <ef_construction> <ef_search> <seed> <device> <out_prefix>
// writes  <out_prefix>.ids.u32 (n: input row -> PointId), <out_prefix>.zero.u32 (n x 2M), <out_prefix>.upper<l>.u32 (n_l x M)
// and     <out_prefix>.meta  (text: n dim M n_layers layer_n... build_seconds).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/instant_distance_b200.h"

static bool write_all(const std::string& path, const void* p, size_t bytes) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes;
    return std::fclose(f) == 0 && ok;
}

int main(int argc, char** argv) {
    if (argc != 10) {
        std::fprintf(stderr, "usage: %s points.f32 n dim M ef_construction ef_search seed device out_prefix\n", argv[0]);
        return 2;
    }
    const uint64_t n = std::strtoull(argv[2], nullptr, 10);
    const uint32_t dim = (uint32_t)std::strtoul(argv[3], nullptr, 10);
    idb_params p;
    if (idb_params_default(&p) != IDB_OK) return 1;
    p.M = (uint32_t)std::strtoul(argv[4], nullptr, 10);
    p.ml = 1.0f / std::log((float)p.M);
    p.ef_construction = (uint32_t)std::strtoul(argv[5], nullptr, 10);
    p.ef_search = (uint32_t)std::strtoul(argv[6], nullptr, 10);
    p.seed = std::strtoull(argv[7], nullptr, 10);
    p.device = std::atoi(argv[8]);
    const std::string out = argv[9];

    std::vector<float> rows((size_t)n * dim);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(rows.data(), 4, rows.size(), f) != rows.size()) {
        std::fprintf(stderr, "cannot read %llu x %u f32 from %s\n", (unsigned long long)n, dim, argv[1]);
        return 1;
    }
    std::fclose(f);

    std::vector<uint32_t> ids(n);
    idb_index* ix = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (idb_build_f32(rows.data(), n, dim, &p, &ix, ids.data()) != IDB_OK) {
        std::fprintf(stderr, "idb_build_f32 failed: %s\n", idb_last_error());
        return 1;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    idb_info info;
    if (idb_index_info(ix, &info) != IDB_OK) return 1;
    bool ok = write_all(out + ".ids.u32", ids.data(), ids.size() * 4);
    {
        std::vector<uint32_t> zero((size_t)n * 2 * p.M);
        ok = ok && idb_index_export_zero(ix, zero.data()) == IDB_OK && write_all(out + ".zero.u32", zero.data(), zero.size() * 4);
    }
    for (uint32_t l = 1; ok && l < info.n_layers; ++l) {
        std::vector<uint32_t> u((size_t)info.layer_n[l] * p.M);
        ok = idb_index_export_upper(ix, l, u.data()) == IDB_OK && write_all(out + ".upper" + std::to_string(l) + ".u32", u.data(), u.size() * 4);
    }
    idb_index_free(ix);
    if (ok) {
        std::string meta = std::to_string(n) + " " + std::to_string(dim) + " " + std::to_string(p.M) + " " + std::to_string(info.n_layers);
        for (uint32_t l = 0; l < info.n_layers; ++l) meta += " " + std::to_string(info.layer_n[l]);
        meta += " " + std::to_string(secs) + "\n";
        ok = write_all(out + ".meta", meta.data(), meta.size());
    }
    if (!ok) {
        std::fprintf(stderr, "export failed: %s\n", idb_last_error());
        return 1;
    }
    std::printf("built %llu x %u, M=%u, %u layers in %.2f s\n", (unsigned long long)n, dim, p.M, info.n_layers, secs);
    return 0;
}
