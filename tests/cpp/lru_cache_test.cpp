// Randomised check of kzero_b200/csrc/selfplay/lru_cache.hpp against a plain std::list + std::unordered_map LRU model.
// Compiled and run by tests/test_host_units.py.  Prints "ok <operations>" or the first mismatch.
#include <cstdio>
#include <list>
#include <unordered_map>

#include "../../kzero_b200/csrc/selfplay/games.hpp"
#include "../../kzero_b200/csrc/selfplay/lru_cache.hpp"

using namespace kzb::selfplay;

struct Model {
    size_t cap;
    std::list<std::pair<uint64_t, float>> order;  // most recent first; the float stands for the stored evaluation
    std::unordered_map<uint64_t, std::list<std::pair<uint64_t, float>>::iterator> map;
    const float* get(uint64_t k) {
        auto it = map.find(k);
        if (it == map.end()) return nullptr;
        order.splice(order.begin(), order, it->second);
        return &it->second->second;
    }
    void put(uint64_t k, float v) {
        if (cap == 0) return;
        auto it = map.find(k);
        if (it != map.end()) {
            it->second->second = v;
            order.splice(order.begin(), order, it->second);
            return;
        }
        order.emplace_front(k, v);
        map[k] = order.begin();
        if (map.size() > cap) {
            map.erase(order.back().first);
            order.pop_back();
        }
    }
};

int run(size_t cap, uint64_t key_space, uint64_t key_stride, int ops, uint64_t seed) {
    LruCache cache(cap);
    Model model{cap, {}, {}};
    Rng rng(seed);
    for (int op = 0; op < ops; op++) {
        // key_stride = 1 << 11 makes every key land in the same bucket of a cap-800 cache (2048 buckets): long chains
        const uint64_t key = (rng.next_u64() % key_space) * key_stride + (key_stride > 1 ? 5 : 0);
        const uint32_t what = rng.gen_range(100);
        if (what < 55) {
            const LruCache::Entry* e = cache.get(key);
            const float* m = model.get(key);
            if ((e != nullptr) != (m != nullptr) || (e && (e->values.value != *m || e->policy.size() != size_t(1 + key % 7) || e->policy[0] != *m))) {
                std::printf("mismatch at op %d: get(%llu) cache %s model %s\n", op, (unsigned long long)key, e ? "hit" : "miss", m ? "hit" : "miss");
                return 1;
            }
        } else if (what < 99) {
            const float v = float(rng.gen_range(1000000));
            if (LruCache::Entry* e = cache.put(key)) {
                e->values.value = v;
                e->policy.assign(size_t(1 + key % 7), v);
            } else if (cap != 0) {
                std::printf("put returned null at op %d\n", op);
                return 1;
            }
            model.put(key, v);
        } else {
            cache.clear();
            model.order.clear();
            model.map.clear();
        }
    }
    return 0;
}

int main() {
    int total = 0;
    const struct { size_t cap; uint64_t space, stride; int ops; } cases[] = {
        {0, 10, 1, 1000}, {1, 5, 1, 20000}, {3, 8, 1, 50000}, {16, 40, 1, 100000}, {800, 3000, 1, 400000},
        {800, 1500, 1ull << 11, 200000}, {64, 200, 1ull << 40, 100000}};
    for (const auto& c : cases) {
        if (run(c.cap, c.space, c.stride, c.ops, 17 + c.cap)) return 1;
        total += c.ops;
    }
    std::printf("ok %d\n", total);
    return 0;
}
