// C++ restatement of the reference's `map` test (instant-distance/tests/all.rs:9-39) through the host mirror header.
// argv[1] == "nodevice": only check that misuse/device errors surface as exceptions (no CUDA device needed).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../instant-distance_b200/cpp/instant_distance.hpp"

using namespace instant_distance;

int main(int argc, char** argv) {
    std::vector<Point> points;
    for (int i = 0; i < 5; ++i) points.push_back(Point{(float)i, (float)i});
    std::vector<std::string> values = {"zero", "one", "two", "three", "four"};
    if (argc > 1 && !std::strcmp(argv[1], "nodevice")) {
        try {
            Builder().seed(1).build(points, values);
            std::puts("FAIL: expected an exception without a device");
            return 1;
        } catch (const Error& e) {
            if (e.status != IDB_ERR_CUDA) { std::printf("FAIL: status %d (%s)\n", (int)e.status, e.what()); return 1; }
            std::printf("ok (no device): %s\n", e.what());
            return 0;
        }
    }
    for (uint64_t seed = 0; seed < 4; ++seed) {
        auto map = Builder().seed(seed).build(points, values);
        Search search;
        auto items = map.search(Point{2.0f, 2.0f}, search);
        if (items.size() != 5) { std::printf("FAIL: %zu items\n", items.size()); return 1; }
        for (size_t i = 0; i < items.size(); ++i) {
            const float d = std::sqrt(items[i].distance);  // the reference test's Point uses sqrt-Euclid; the engine is squared-L2
            const std::string& v = *items[i].value;
            bool ok = (i == 0 && d == 0.0f && v == "two") || ((i == 1 || i == 2) && d == 1.4142135f && (v == "one" || v == "three")) ||
                      ((i == 3 || i == 4) && d == 2.828427f && (v == "zero" || v == "four"));
            if (!ok) { std::printf("FAIL: seed %llu item %zu d=%.9g v=%s\n", (unsigned long long)seed, i, d, v.c_str()); return 1; }
        }
    }
    auto [hnsw, ids] = Hnsw::builder().seed(7).ef_search(3).build_hnsw(points);
    Search s;
    auto r = hnsw.search(points[4], s);
    if (r.size() != 3 || r[0].distance != 0.f || !(hnsw[r[0].pid].v == points[4].v) || !ids[4].is_valid() || ids[4].raw != r[0].pid.raw) {
        std::puts("FAIL: build_hnsw/search");
        return 1;
    }
    if (Point{0.f, 3.f}.distance(Point{4.f, 0.f}) != 25.0f) { std::puts("FAIL: distance"); return 1; }
    std::puts("ok: map test through instant_distance.hpp");
    return 0;
}
