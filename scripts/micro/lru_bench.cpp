// Host micro-benchmark of the per-game evaluation cache under the generator's access pattern: `games` caches visited round-robin
// (by the time a cache is visited again its lines have left the CPU caches), per visit 16 x (lookup that misses, insert with a
// ~31-float policy) plus a few lookups that hit.
//   g++ -O3 -std=c++17 -o /tmp/lru_bench scripts/micro/lru_bench.cpp && /tmp/lru_bench [games] [seconds]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "../../kzero_b200/csrc/selfplay/games.hpp"
#include "../../kzero_b200/csrc/selfplay/lru_cache.hpp"

using namespace kzb::selfplay;

int main(int argc, char** argv) {
    const int games = argc > 1 ? atoi(argv[1]) : 384;
    const double seconds = argc > 2 ? atof(argv[2]) : 3.0;
    std::vector<std::unique_ptr<LruCache>> caches;
    for (int g = 0; g < games; g++) caches.push_back(std::make_unique<LruCache>(800));
    std::vector<uint64_t> next_key(size_t(games), 1);
    float policy[64];
    for (int i = 0; i < 64; i++) policy[i] = float(i) * 0.01f;
    uint64_t ops = 0, hits = 0;
    double sink = 0;
    const auto t0 = std::chrono::steady_clock::now();
    double el = 0;
    while (el < seconds) {
        for (int g = 0; g < games; g++) {
            LruCache& c = *caches[size_t(g)];
            uint64_t& k = next_key[size_t(g)];
            for (int i = 0; i < 16; i++) {
                const uint64_t key = splitmix64(uint64_t(g) * 1000003ull + k);
                c.prefetch(key);
                if (c.get(key)) hits++;
                if (LruCache::Entry* e = c.put(key)) {
                    e->values = ValuesPov{0.1f, 0.2f, 0.3f, 0.4f, 0.5f};
                    const size_t n = 20 + size_t(key % 25);
                    e->policy.assign(policy, policy + n);
                }
                k++;
                ops++;
            }
            for (int i = 0; i < 2; i++) {  // revisits of recent positions (the search's transpositions): hits
                const uint64_t key = splitmix64(uint64_t(g) * 1000003ull + k - 1 - uint64_t(i) * 37);
                if (const LruCache::Entry* e = c.get(key)) hits++, sink += e->policy[0];
            }
        }
        el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    std::printf("games %d: %.1f ns per lookup+insert  (%llu inserts, %llu hits, %g)\n", games, el * 1e9 / double(ops), (unsigned long long)ops,
                (unsigned long long)hits, sink);
    return 0;
}
