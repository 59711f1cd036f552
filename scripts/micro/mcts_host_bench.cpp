// Host-side MCTS micro-benchmark: `slots` concurrent trees visited round-robin (the generator thread's access pattern:
// by the time a tree is visited again its nodes have left the cache), pseudo-network answers applied one round later.
//   g++ -O3 -std=c++17 -o scripts/micro/mcts_host_bench scripts/micro/mcts_host_bench.cpp && scripts/micro/mcts_host_bench [slots] [visits] [seconds]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <x86intrin.h>

#include "../../kzero_b200/csrc/selfplay/games.hpp"
#include "../../kzero_b200/csrc/selfplay/mcts.hpp"

using namespace kzb::selfplay;
using Game = SynthChess;

struct Slot {
    Game board;
    std::unique_ptr<Tree<Game>> tree;
    Rng rng;
    std::vector<Request<Game>> requests;
    Descent<Game> descent;
    int terminal = 0;
    uint64_t seed;
    explicit Slot(uint64_t s) : board(Game::start(s)), rng(s), seed(s) { reset(); }
    void reset() {
        tree = std::make_unique<Tree<Game>>(board);
        tree->reserve(800 * 48 + 64, 800 * 2 + 64);
    }
};

int main(int argc, char** argv) {
    const int slots_n = argc > 1 ? atoi(argv[1]) : 48, visits = argc > 2 ? atoi(argv[2]) : 800;
    const double seconds = argc > 3 ? atof(argv[3]) : 4.0;
    const int group_n = argc > 4 ? atoi(argv[4]) : 1;  // trees whose descents are interleaved
    std::vector<Slot*> group;
    SearchSettings settings;
    std::vector<std::unique_ptr<Slot>> slots;
    for (int i = 0; i < slots_n; i++) slots.push_back(std::make_unique<Slot>(uint64_t(i) * 7919 + 1));
    std::vector<uint32_t> scratch;
    std::vector<float> policy;
    uint64_t nodes = 0, cyc_gather = 0, cyc_apply = 0, gathers = 0;
    // interleaved mode: the descents of the grouped trees advance one level at a time, round-robin
    auto flush_group = [&] {
        if (group.empty()) return;
        const uint64_t c1 = __rdtsc();
        size_t active = group.size();
        while (active) {
            for (size_t g = 0; g < group.size(); g++) {
                Slot* gs = group[g];
                if (!gs) continue;
                Request<Game> req;
                const StepResult r = descent_step(*gs->tree, settings, gs->rng, gs->descent, req, scratch);
                if (r == StepResult::kDescend) continue;
                gathers++;
                if (r == StepResult::kRequest) gs->requests.push_back(std::move(req));
                else gs->terminal++;
                if (int(gs->requests.size()) < 16 && gs->terminal < 16) gs->descent.begin(*gs->tree);
                else group[g] = nullptr, active--;
            }
        }
        group.clear();
        cyc_gather += __rdtsc() - c1;
    };
    const auto t0 = std::chrono::steady_clock::now();
    double el = 0;
    while (el < seconds) {
        for (auto& sp : slots) {
            Slot& s = *sp;
            Tree<Game>& tree = *s.tree;
            uint64_t c0 = __rdtsc();
            for (auto& req : s.requests) {  // answers of the previous round
                const uint64_t h = req.board.hash();
                const size_t n = size_t(tree.nodes[size_t(req.node)].child_count);
                policy.resize(n);
                float sum = 0;
                for (size_t k = 0; k < n; k++) {
                    const float u = float((splitmix64(h + k + 1) >> 40) % 1000 + 1) * 1e-3f;
                    policy[k] = u * u * u * u;
                    sum += policy[k];
                }
                for (auto& p : policy) p /= sum;
                const float v = float(int((splitmix64(h ^ 0xABCDull) >> 40) % 2001) - 1000) / 1000.0f;
                const ValuesPov vals{v, (1 + v) * 0.4f, 0.2f, (1 - v) * 0.4f, float((h >> 50) % 50)};
                zero_step_apply(tree, req.node, req.board.next_player(), vals, policy.data(), n);
            }
            nodes += s.requests.size();
            s.requests.clear();
            uint64_t c1 = __rdtsc();
            cyc_apply += c1 - c0;
            if (tree.root_visits() >= uint64_t(visits)) {
                tree.policy(policy);
                size_t best = 0;
                for (size_t i = 1; i < policy.size(); i++)
                    if (policy[i] > policy[best]) best = i;
                s.board.play(tree.move_of(tree.root().child_start, tree.root().child_count, int(best)));
                if (s.board.done()) s.board = Game::start(++s.seed * 104729);
                s.reset();
                continue;
            }
            if (group_n <= 1) {
                int terminal = 0;
                while (int(s.requests.size()) < 16 && terminal < 16) {
                    Request<Game> req;
                    gathers++;
                    if (zero_step_gather(tree, settings, s.rng, req, scratch)) s.requests.push_back(std::move(req));
                    else terminal++;
                }
                cyc_gather += __rdtsc() - c1;
                continue;
            }
            s.terminal = 0;
            s.descent.begin(tree);
            group.push_back(&s);
            if (int(group.size()) == group_n) flush_group();
        }
        flush_group();
        el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    printf("slots %d: %.0f nodes/s  (%.2f us/node)  gather %.0f cycles/node  apply %.0f cycles/node  gathers/node %.2f\n", slots_n, nodes / el,
           1e6 * el / nodes, double(cyc_gather) / nodes, double(cyc_apply) / nodes, double(gathers) / nodes);
    return 0;
}
