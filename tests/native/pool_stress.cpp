#include "mmrs_pool.hpp"
#include <cstdio>
#include <stdexcept>
#include <numeric>
using namespace mmrs;
int main() {
    setenv("MMRS_HOST_THREADS", "8", 1);
    // nested + concurrent submitters
    std::vector<std::thread> subs;
    std::atomic<long long> total{0};
    for (int s = 0; s < 6; ++s)
        subs.emplace_back([&] {
            for (int rep = 0; rep < 200; ++rep) {
                std::vector<long long> out(64, 0);
                parallel_for(out.size(), [&](size_t i) {
                    std::vector<long long> inner(37, 0);
                    parallel_for(inner.size(), [&](size_t k) { inner[k] = (long long)(i * 1000 + k); }, 4);
                    out[i] = std::accumulate(inner.begin(), inner.end(), 0ll);
                });
                long long sum = std::accumulate(out.begin(), out.end(), 0ll);
                total += sum;
            }
        });
    for (auto& t : subs) t.join();
    long long want = 0;
    for (int i = 0; i < 64; ++i) for (int k = 0; k < 37; ++k) want += i * 1000 + k;
    std::printf("total %lld want %lld\n", total.load(), want * 6 * 200);
    // lowest-index exception
    int ok = 0;
    for (int rep = 0; rep < 200; ++rep) {
        try {
            parallel_for(100, [&](size_t i) { if (i == 17 || i == 40 || i == 93) throw std::runtime_error(std::to_string(i)); });
        } catch (const std::runtime_error& e) { if (std::string(e.what()) == "17") ++ok; }
    }
    std::printf("lowest-index exceptions %d / 200\n", ok);
    return (total.load() == want * 6 * 200 && ok == 200) ? 0 : 1;
}
