// Self-play driver: generator threads running many concurrent MCTS games + executor threads that batch their
// evaluation requests into the B200 evaluator (BASELINE.json configs[3]).
//
// Restates the structure of the reference's self-play server for the AlphaZero path:
//   executor  : batched_executor_loop, rust/kz-selfplay/src/server/executor.rs:27-146 -- collect jobs until the batch is
//               full or RunCondition::JobCount(gpu_batch / search_batch) jobs are queued (:240-253), evaluate, scatter
//               the results back to the jobs in order (:278-301)
//   generator : generate_simulation / build_tree / apply_eval, rust/kz-selfplay/src/server/generator_alphazero.rs:70-260 --
//               per move: gather up to search_batch requests (virtual loss), serve LRU-cache hits immediately, send the
//               rest as one job, apply the answers (policy temperature, Dirichlet noise at the root), until the root has
//               `visits` visits; then pick a move (MoveSelector) and start a new tree
//   topology  : cpu_threads generator threads and gpu_threads executor instances per device, concurrent_games =
//               ceil((gpu_threads + 1) * gpu_batch / search_batch), rust/kz-selfplay/src/server/server_alphazero.rs:47-121
// Differences that are deliberate: generator threads are plain threads that round-robin their games instead of async
// tasks; boards are ENCODED ON THE GENERATOR THREADS into the packed (bits, scalars, legal-index) record, so the
// executor thread only concatenates records and calls the evaluator (the reference encodes f32 planes on the executor
// thread, network/cudnn.rs:62-64); game records are written by whichever generator thread finishes a game (record_writer.hpp,
// row N2) instead of a collector thread.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <list>
#include <memory>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <thread>
#include <unordered_map>
#if defined(__x86_64__)
#include <x86intrin.h>
#endif

#include "../../../include/kzb200.h"
#include "../executor.hpp"
#include "chess_game.hpp"
#include "games.hpp"
#include "lru_cache.hpp"
#include "mcts.hpp"
#include "record_writer.hpp"

#define KZB_API extern "C" __attribute__((visibility("default")))

namespace kzb {
namespace selfplay {
namespace {

std::atomic<bool> g_stop_requested{false};
std::atomic<bool> g_interrupt_requested{false};  // kzb_selfplay_request_interrupt: session runs return, their record files stay open

// KZB_SP_PROFILE=1: per-section cycle counts of the generator threads, printed when a run ends (host tuning aid)
enum Section { kSecApply, kSecMove, kSecGather, kSecEncode, kSecQueue, kSecCount };
const char* const kSectionNames[kSecCount] = {"apply answers", "pick move / new tree", "gather", "encode job", "enqueue"};
struct SectionClock {
    bool on = false;
    uint64_t cycles[kSecCount] = {}, last = 0;
    static uint64_t now() {
#if defined(__x86_64__)
        return __rdtsc();
#else
        return uint64_t(std::chrono::steady_clock::now().time_since_epoch().count());
#endif
    }
    void start() { if (on) last = now(); }
    void lap(Section s) {
        if (!on) return;
        const uint64_t t = now();
        cycles[s] += t - last;
        last = t;
    }
};
std::mutex g_profile_mu;
uint64_t g_profile_cycles[kSecCount];
double g_profile_cpu_s[2];  // CPU seconds of the generator / executor threads
double thread_cpu_seconds() {
    timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}

struct Eval {
    ValuesPov values;
    std::vector<float> policy;
};

SearchSettings search_settings(const kzb_selfplay_config& c) {
    SearchSettings s;
    s.weights = {c.exploration_weight, c.moves_left_weight, c.moves_left_clip, c.moves_left_sharpness};
    s.q_mode = {c.q_mode_wdl != 0, c.draw_score};
    s.fpu_root = {c.fpu_root_relative != 0, c.fpu_root};
    s.fpu_child = {c.fpu_child_relative != 0, c.fpu_child};
    s.virtual_loss = c.virtual_loss;
    return s;
}

// one evaluation job = the requests of one gather round of one game, already encoded (job_channel.rs:9-22)
struct Job {
    int n = 0;
    std::vector<uint8_t> bits;
    std::vector<float> scalars;
    std::vector<uint32_t> mv_idx, mv_off;  // local CSR, mv_off[0] = 0
    std::vector<float> values, probs;      // filled by the executor
    std::atomic<int> done{0};
    int owner = 0;  // generator thread to wake
};

struct Shared {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job*> queue;
    size_t queued_positions = 0;
    int executors_busy = 0;  // executors between taking a batch and having its answers (guarded by mu)
    std::atomic<bool> stop{false};
    size_t job_count = 1;  // RunCondition::JobCount
    std::string error;
    // per generator thread wake-up
    std::vector<std::unique_ptr<std::mutex>> gen_mu;
    std::vector<std::unique_ptr<std::condition_variable>> gen_cv;
    std::vector<uint64_t> gen_signal;  // answers delivered to each generator thread so far (guarded by its gen_mu)
    // finished games go to one writer, the collector's role (collector.rs:59-85)
    std::mutex writer_mu;
    std::unique_ptr<RecordWriter> writer;
    // statistics (collector.rs:172-191: real / cached evals)
    std::atomic<uint64_t> real_evals{0}, cached_evals{0}, batches{0}, games{0}, moves{0}, root_visits{0}, max_batch_seen{0},
        potential_evals{0};
};

template <typename Game>
struct Slot {
    Game board;
    std::unique_ptr<Tree<Game>> tree;
    LruCache cache;
    Rng rng;
    uint32_t move_count = 0;
    uint64_t target_visits = 0;  // 0: not drawn yet for this move
    bool is_full_search = true;
    uint64_t next_seed;
    bool waiting = false;
    std::vector<Request<Game>> requests;
    Job job;
    RecordedGame record;   // only filled when records are written
    Eval root_net_eval;    // the network's own evaluation of the root (generator_alphazero.rs:226-229)
    Slot(uint64_t seed, size_t cache_size, size_t visits, uint32_t max_game_length)
        : board(Game::start(seed)), cache(cache_size), rng(seed ^ 0x5EEDull), next_seed(seed + 0x1000) {
        board.max_moves = max_game_length;
        tree = std::make_unique<Tree<Game>>(board);
        tree->reserve(visits * 48 + 64, visits * 2 + 64);
    }
};

// generator_alphazero.rs:217-245.  `policy` is read-only (a slice of the job's answer or a cache entry); the copy that
// temperature and noise need goes through `tmp`, and only when they actually change something.
template <typename Game>
void apply_eval(Slot<Game>& slot, const Request<Game>& req, const ValuesPov& values, const float* policy, size_t n,
                const kzb_selfplay_config& c, std::vector<float>& tmp) {
    const bool root = req.is_root();
    if (root) {  // the network's own answer, before temperature and noise
        slot.root_net_eval.values = values;
        slot.root_net_eval.policy.assign(policy, policy + n);
    }
    const float temperature = root ? c.policy_temperature_root : c.policy_temperature_child;
    if (temperature != 1.0f || root) {
        tmp.assign(policy, policy + n);
        policy_softmax_temperature_in_place(tmp.data(), n, temperature);
        if (root) add_dirichlet_noise(tmp.data(), n, c.dirichlet_alpha, c.dirichlet_eps, slot.rng);
        policy = tmp.data();
    }
    zero_step_apply(*slot.tree, req.node, req.board.next_player(), values, policy, n);
}

template <typename Game>
void encode_record(const Game& b, const GameShape& shape, RecordedPosition& rp) {
    rp.bits.assign(size_t(shape.bits_bytes()), 0);
    rp.scalars.assign(size_t(shape.scalar_count), 0.0f);
    b.encode(rp.bits.data(), rp.scalars.data());
    rp.next_player = b.next_player();
}

template <typename Game>
void generator_main(int tid, std::vector<std::unique_ptr<Slot<Game>>>& slots, Shared& sh, const kzb_selfplay_config& c) {
    const SearchSettings settings = search_settings(c);
    const GameShape shape = Game::shape();
    const int bits_bytes = shape.bits_bytes();
    std::vector<uint32_t> scratch;
    std::vector<float> policy_tmp;
    SectionClock clock;
    clock.on = std::getenv("KZB_SP_PROFILE") != nullptr;
    const char* spin_env = std::getenv("KZB_SP_SPIN_US");
    const long spin_us = spin_env ? std::atol(spin_env) : 0;
    uint64_t seen_signal = 0;
    if (std::getenv("KZB_SP_PIN_GENERATORS")) {  // generator t stays on the t-th CPU this process may use
        cpu_set_t allowed, one;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            int seen = 0, want = tid % std::max(1, CPU_COUNT(&allowed));
            for (int cpu = 0; cpu < CPU_SETSIZE; cpu++)
                if (CPU_ISSET(cpu, &allowed) && seen++ == want) {
                    CPU_ZERO(&one);
                    CPU_SET(cpu, &one);
                    pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
                    break;
                }
        }
    }
    try {
        while (!sh.stop.load(std::memory_order_relaxed)) {
            bool progressed = false;
            for (auto& sp : slots) {
                Slot<Game>& slot = *sp;
                if (slot.waiting) {
                    if (!slot.job.done.load(std::memory_order_acquire)) continue;
                    clock.start();
                    // the lines the loop below will write were last touched a GPU round trip ago: ask for all of them
                    // up front so that their misses overlap instead of queueing up one request at a time
                    prefetch_span(slot.job.probs.data(), slot.job.probs.size() * 4);  // written by the executor's core
                    prefetch_span(slot.job.values.data(), slot.job.values.size() * 4);
                    for (int i = 0; i < slot.job.n; i++) {
                        const Request<Game>& req = slot.requests[size_t(i)];
                        const Node& node = slot.tree->nodes[size_t(req.node)];
                        __builtin_prefetch(&node, 1);
                        slot.cache.prefetch(req.board.hash());
                        prefetch_span(slot.tree->policy_of(req.child_start), size_t(req.child_count) * 4);
                    }
                    // answers are back: cache + apply in request order (generator_alphazero.rs:206-210)
                    for (int i = 0; i < slot.job.n; i++) {
                        const float* v = slot.job.values.data() + size_t(i) * 5;
                        const ValuesPov values{v[0], v[1], v[2], v[3], v[4]};
                        const float* probs = slot.job.probs.data() + slot.job.mv_off[size_t(i)];
                        const size_t count = slot.job.mv_off[size_t(i) + 1] - slot.job.mv_off[size_t(i)];
                        if (LruCache::Entry* e = slot.cache.put(slot.requests[size_t(i)].board.hash())) {
                            e->values = values;
                            e->policy.assign(probs, probs + count);
                        }
                        apply_eval(slot, slot.requests[size_t(i)], values, probs, count, c, policy_tmp);
                    }
                    slot.waiting = false;
                    clock.lap(kSecApply);
                }
                progressed = true;
                clock.start();
                Tree<Game>& tree = *slot.tree;
                if (slot.target_visits == 0) {  // generator_alphazero.rs:88-94
                    slot.is_full_search = c.full_search_prob >= 1.0f || slot.rng.gen_bool(c.full_search_prob);
                    slot.target_visits = uint64_t(slot.is_full_search ? c.visits : std::max(1, c.part_iterations));
                }
                if (tree.root_visits() >= slot.target_visits) {
                    // pick and play a move (generator_alphazero.rs:108-130)
                    std::vector<float> policy;
                    tree.policy(policy);
                    const size_t pick = select_move(policy.data(), policy.size(), slot.move_count, c.temperature, uint32_t(c.zero_temp_move_count), slot.rng);
                    const size_t c0 = size_t(tree.root().child_start), cn = size_t(tree.root().child_count);
                    const uint32_t mv = tree.move_of(int(c0), int(cn), int(pick));
                    sh.root_visits.fetch_add(tree.root_visits(), std::memory_order_relaxed);
                    sh.moves.fetch_add(1, std::memory_order_relaxed);
                    if (sh.writer) {  // Position, generator_alphazero.rs:115-123
                        RecordedPosition rp;
                        encode_record(slot.board, shape, rp);
                        for (size_t k = 0; k < cn; k++) rp.indices.push_back(slot.board.move_to_index(tree.move_of(int(c0), int(cn), int(k))));
                        rp.played_index = slot.board.move_to_index(mv);
                        rp.is_full_search = slot.is_full_search;
                        rp.zero_visits = tree.root_visits();
                        rp.zero_values = pov(tree.root_values(), slot.board.next_player());  // Tree::values, tree.rs:95-98
                        rp.zero_policy = policy;
                        rp.net_values = slot.root_net_eval.values;
                        rp.net_policy = slot.root_net_eval.policy;
                        slot.record.positions.push_back(std::move(rp));
                    }
                    slot.board.play(mv);
                    slot.move_count++;
                    slot.target_visits = 0;
                    slot.board.max_moves = uint32_t(c.max_game_length);  // Settings::max_game_length may change between generations
                    if (slot.board.done()) {  // the game's own end, or the length cap (a draw: MaxMovesBoard)
                        if (sh.writer) {
                            encode_record(slot.board, shape, slot.record.final_position);
                            slot.record.final_done = slot.board.inner.done();
                            slot.record.outcome = slot.board.outcome();
                            {
                                std::lock_guard<std::mutex> lk(sh.writer_mu);
                                sh.writer->append(slot.record);
                            }
                            slot.record = RecordedGame();
                        }
                        sh.games.fetch_add(1, std::memory_order_relaxed);
                        slot.board = Game::start(slot.next_seed++);
                        slot.board.max_moves = uint32_t(c.max_game_length);
                        slot.move_count = 0;
                        slot.cache.clear();  // a new cache for every game (generator_alphazero.rs:77-79)
                    }
                    slot.tree->reset(slot.board);
                    slot.tree->reserve(size_t(c.visits) * 48 + 64, size_t(c.visits) * 2 + 64);  // no-op unless `visits` grew
                    clock.lap(kSecMove);
                    continue;
                }
                // collect a batch of requests (generator_alphazero.rs:165-201)
                slot.requests.clear();
                if (slot.requests.capacity() <= size_t(c.search_batch)) slot.requests.reserve(size_t(c.search_batch) + 1);
                int terminal_gathers = 0;
                uint64_t cached = 0;
                while (int(slot.requests.size()) < c.search_batch && terminal_gathers < c.search_batch) {
                    // gathered in place: a request holds a whole board, and only requests the cache cannot answer stay in the list
                    slot.requests.emplace_back();
                    Request<Game>& req = slot.requests.back();
                    if (zero_step_gather(tree, settings, slot.rng, req, scratch, [&](const Game& b) { slot.cache.prefetch(b.hash()); })) {
                        if (const LruCache::Entry* hit = slot.cache.get(req.board.hash())) {
                            cached++;
                            apply_eval(slot, req, hit->values, hit->policy.data(), hit->policy.size(), c, policy_tmp);
                            slot.requests.pop_back();
                        }
                    } else {
                        terminal_gathers++;
                        slot.requests.pop_back();
                    }
                }
                if (cached) sh.cached_evals.fetch_add(cached, std::memory_order_relaxed);
                clock.lap(kSecGather);
                if (slot.requests.empty()) continue;
                // encode the requests into one job
                Job& job = slot.job;
                job.n = int(slot.requests.size());
                job.owner = tid;
                job.bits.resize(size_t(job.n) * bits_bytes);
                job.scalars.resize(size_t(job.n) * shape.scalar_count);
                job.mv_off.resize(size_t(job.n) + 1);
                job.mv_off[0] = 0;
                for (int i = 0; i < job.n; i++)
                    job.mv_off[size_t(i) + 1] = job.mv_off[size_t(i)] + uint32_t(tree.nodes[size_t(slot.requests[size_t(i)].node)].child_count);
                job.mv_idx.resize(job.mv_off[size_t(job.n)]);
                for (int i = 0; i < job.n; i++) {
                    const Game& b = slot.requests[size_t(i)].board;
                    b.encode(job.bits.data() + size_t(i) * bits_bytes, job.scalars.data() + size_t(i) * shape.scalar_count);
                    // the node's children were created from available_moves() in order (step.rs:89-97): their moves ARE the legal list
                    const int node = slot.requests[size_t(i)].node;
                    const int c0 = tree.nodes[size_t(node)].child_start;
                    uint32_t* idx = job.mv_idx.data() + job.mv_off[size_t(i)];
                    const size_t cn = size_t(tree.nodes[size_t(node)].child_count);
                    for (size_t k = 0; k < cn; k++) idx[k] = b.move_to_index(tree.move_of(c0, int(cn), int(k)));
                }
                job.values.resize(size_t(job.n) * 5);
                job.probs.resize(job.mv_idx.size());
                job.done.store(0, std::memory_order_relaxed);
                slot.waiting = true;
                clock.lap(kSecEncode);
                bool wake;
                {
                    std::lock_guard<std::mutex> lk(sh.mu);
                    sh.queue.push_back(&job);
                    sh.queued_positions += size_t(job.n);
                    // only wake an executor when its run condition has just become true (executor.rs:240-253)
                    wake = sh.queued_positions >= size_t(c.gpu_batch) || sh.queue.size() >= sh.job_count;
                }
                if (wake) sh.cv.notify_one();
                clock.lap(kSecQueue);
            }
            if (!progressed) {
                // nothing is ready: spin for a moment (a sleeping thread of a busy virtual machine comes back late), then
                // sleep until an executor has delivered answers to this thread
                bool arrived = false;
                if (spin_us > 0) {
                    const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us);
                    while (!arrived && std::chrono::steady_clock::now() < until) {
                        for (int i = 0; i < 64 && !arrived; i++) {
                            for (auto& sp : slots) arrived |= sp->waiting && sp->job.done.load(std::memory_order_acquire) != 0;
#if defined(__x86_64__)
                            _mm_pause();
#endif
                        }
                    }
                }
                if (!arrived) {
                    std::unique_lock<std::mutex> lk(*sh.gen_mu[size_t(tid)]);
                    sh.gen_cv[size_t(tid)]->wait_for(lk, std::chrono::microseconds(100), [&] { return sh.gen_signal[size_t(tid)] != seen_signal; });
                    seen_signal = sh.gen_signal[size_t(tid)];
                }
            }
        }
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(sh.mu);
        if (sh.error.empty()) sh.error = std::string("generator thread: ") + e.what();
        sh.stop.store(true);
        sh.cv.notify_all();
    }
    if (clock.on) {
        std::lock_guard<std::mutex> lk(g_profile_mu);
        for (int i = 0; i < kSecCount; i++) g_profile_cycles[i] += clock.cycles[i];
        g_profile_cpu_s[0] += thread_cpu_seconds();
    }
}

// `net == nullptr`: the reference's DummyNetwork (uniform wdl and policy, rust/kz-core/src/network/dummy.rs:44-60, what the
// server uses after a UseDummyNetwork command) -- runs the whole driver without a GPU
void executor_main(Net* net, Shared& sh, const kzb_selfplay_config& c, const GameShape shape) {
    const size_t job_count = sh.job_count;
    const int bits_bytes = shape.bits_bytes();
    std::vector<uint8_t> bits(size_t(c.gpu_batch) * bits_bytes);
    std::vector<char> owner_woken(sh.gen_cv.size(), 0);
    std::vector<float> scalars(size_t(c.gpu_batch) * shape.scalar_count), values(size_t(c.gpu_batch) * 5), probs;
    std::vector<uint32_t> mv_idx, mv_off;
    std::vector<Job*> jobs;
    // KZB_SP_DUMMY_LATENCY_US: make the dummy network take as long as a GPU evaluation would, so that a host-only run batches
    // like a real one (profiling aid)
    const char* latency_env = std::getenv("KZB_SP_DUMMY_LATENCY_US");
    const long dummy_latency_us = latency_env ? std::atol(latency_env) : 0;
    const bool dummy_serial = std::getenv("KZB_SP_DUMMY_SERIAL") != nullptr;
    try {
        while (true) {
            jobs.clear();
            size_t n = 0;
            {
                std::unique_lock<std::mutex> lk(sh.mu);
                // executor.rs:240-253 should_eval: a full batch, or JobCount jobs queued.  A partial batch is taken only
                // when no executor has work on the GPU -- then waiting buys nothing, and the last few games of a run
                // (fewer than JobCount producers left) still get answers; while the GPU is busy a partial batch would
                // only spend a launch on fewer positions.
                while (true) {
                    const bool full = sh.cv.wait_for(lk, std::chrono::microseconds(200), [&] {
                        return sh.stop.load() || sh.queued_positions >= size_t(c.gpu_batch) || sh.queue.size() >= job_count;
                    });
                    if (full || (!sh.queue.empty() && sh.executors_busy == 0)) break;
                }
                if (sh.stop.load()) break;
                while (!sh.queue.empty() && n + size_t(sh.queue.front()->n) <= size_t(c.gpu_batch)) {
                    Job* j = sh.queue.front();
                    sh.queue.pop_front();
                    sh.queued_positions -= size_t(j->n);
                    n += size_t(j->n);
                    jobs.push_back(j);
                }
                if (!jobs.empty()) sh.executors_busy++;
            }
            if (jobs.empty()) continue;
            mv_off.assign(1, 0);
            mv_idx.clear();
            size_t row = 0;
            for (Job* j : jobs) {
                std::memcpy(bits.data() + row * bits_bytes, j->bits.data(), j->bits.size());
                std::memcpy(scalars.data() + row * shape.scalar_count, j->scalars.data(), j->scalars.size() * 4);
                const uint32_t base = uint32_t(mv_idx.size());
                mv_idx.insert(mv_idx.end(), j->mv_idx.begin(), j->mv_idx.end());
                for (int i = 1; i <= j->n; i++) mv_off.push_back(base + j->mv_off[size_t(i)]);
                row += size_t(j->n);
            }
            probs.resize(std::max<size_t>(mv_idx.size(), 1));
            if (net) {
                net->eval_packed(bits.data(), scalars.data(), int(n), mv_idx.data(), mv_off.data(), values.data(), probs.data());
            } else {
                if (dummy_latency_us > 0) {
                    // KZB_SP_DUMMY_SERIAL: one batch at a time, like executors sharing one GPU
                    static std::mutex one_gpu;
                    std::unique_lock<std::mutex> gpu(one_gpu, std::defer_lock);
                    if (dummy_serial) gpu.lock();
                    std::this_thread::sleep_for(std::chrono::microseconds(dummy_latency_us));
                }
                for (size_t i = 0; i < n; i++) {
                    const uint32_t cnt = mv_off[i + 1] - mv_off[i];
                    float* p = probs.data() + mv_off[i];
                    if (c.dummy_network == 2) {  // a deterministic pseudo-network: sharp, position-dependent answers
                        uint64_t h = 0xCBF29CE484222325ull;
                        for (int k = 0; k < bits_bytes; k += 8) {
                            uint64_t w = 0;
                            std::memcpy(&w, bits.data() + i * bits_bytes + k, size_t(std::min(8, bits_bytes - k)));
                            h = splitmix64(h ^ w);
                        }
                        float sum = 0.0f;
                        for (uint32_t k = 0; k < cnt; k++) {
                            const float u = float((splitmix64(h + k + 1) >> 40) % 1000 + 1) * 1e-3f;
                            p[k] = u * u * u * u;
                            sum += p[k];
                        }
                        for (uint32_t k = 0; k < cnt; k++) p[k] /= sum;
                        const float val = float(int((splitmix64(h ^ 0xABCDull) >> 40) % 2001) - 1000) / 1000.0f;
                        const float v[5] = {val, (1.0f + val) * 0.4f, 0.2f, (1.0f - val) * 0.4f, float((h >> 50) % 50)};
                        std::memcpy(values.data() + i * 5, v, sizeof(v));
                        continue;
                    }
                    const float v[5] = {0.0f, 1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f, 0.0f};
                    std::memcpy(values.data() + i * 5, v, sizeof(v));
                    for (uint32_t k = 0; k < cnt; k++) p[k] = 1.0f / float(cnt);
                }
            }
            {
                std::lock_guard<std::mutex> lk(sh.mu);
                sh.executors_busy--;
            }
            row = 0;
            for (Job* j : jobs) {
                std::memcpy(j->values.data(), values.data() + row * 5, size_t(j->n) * 5 * 4);
                const uint32_t base = mv_off[row];
                std::memcpy(j->probs.data(), probs.data() + base, j->probs.size() * 4);
                row += size_t(j->n);
                owner_woken[size_t(j->owner)] = 1;
                j->done.store(1, std::memory_order_release);
            }
            for (size_t o = 0; o < owner_woken.size(); o++)  // one wake-up per generator thread and batch
                if (owner_woken[o]) {
                    owner_woken[o] = 0;
                    {
                        std::lock_guard<std::mutex> lk(*sh.gen_mu[o]);  // no lost wake-up: the sleeper re-checks gen_signal
                        sh.gen_signal[o]++;
                    }
                    sh.gen_cv[o]->notify_one();
                }
            sh.real_evals.fetch_add(n, std::memory_order_relaxed);
            sh.potential_evals.fetch_add(uint64_t(c.gpu_batch), std::memory_order_relaxed);
            sh.batches.fetch_add(1, std::memory_order_relaxed);
            uint64_t prev = sh.max_batch_seen.load();
            while (prev < n && !sh.max_batch_seen.compare_exchange_weak(prev, n)) {
            }
        }
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(sh.mu);
        if (sh.error.empty()) sh.error = std::string("executor thread: ") + e.what();
        sh.stop.store(true);
        sh.cv.notify_all();
    }
    if (std::getenv("KZB_SP_PROFILE")) {
        std::lock_guard<std::mutex> lk(g_profile_mu);
        g_profile_cpu_s[1] += thread_cpu_seconds();
    }
}

template <typename Game>
using SlotTable = std::vector<std::vector<std::unique_ptr<Slot<Game>>>>;  // [generator thread][its games]

// `kept`: the games of a session (kzb_selfplay_session_*).  Empty on the first run: filled here; afterwards every game continues
// where the previous run left it -- board, tree, per-game cache and the positions recorded so far -- like the reference's generators,
// which run across file boundaries (collector.rs:59-116 only rotates the output file).  nullptr: games live for this run only.
// a record file that an interrupted session run left open, to be continued by the session's next run with the same prefix
struct OpenRecordFile {
    std::unique_ptr<RecordWriter> writer;
    std::string prefix;
    void finish() {
        if (writer) writer->finish();
        writer.reset();
        prefix.clear();
    }
};

template <typename Game>
void run_selfplay(int device, const void* onnx, size_t len, int precision, const kzb_selfplay_config& c, kzb_selfplay_stats& out,
                  SlotTable<Game>* kept = nullptr, OpenRecordFile* open_file = nullptr) {
    const GameShape shape = Game::shape();
    if (c.visits < 1 || c.search_batch < 1 || c.gpu_batch < c.search_batch || c.cpu_threads < 1 || c.gpu_threads < 1)
        throw std::runtime_error("selfplay config: need visits >= 1, 1 <= search_batch <= gpu_batch, cpu_threads >= 1, gpu_threads >= 1");
    std::vector<std::unique_ptr<Net>> nets;
    for (int i = 0; i < c.gpu_threads && !c.dummy_network; i++) {
        nets.push_back(std::make_unique<Net>(device, onnx, len, c.gpu_batch, precision));
        nets.back()->bind_mapper(shape.scalar_count, shape.bool_channels, shape.board, shape.board, shape.policy_len);
        nets.back()->set_blocking_sync(c.executor_blocking_sync != 0);
    }
    // server_alphazero.rs:47: concurrent_games = ceil((gpu_threads + 1) * gpu_batch / search_batch)
    int games = c.concurrent_games > 0 ? c.concurrent_games : ((c.gpu_threads + 1) * c.gpu_batch + c.search_batch - 1) / c.search_batch;
    Shared sh;
    const std::string prefix = c.output_prefix ? c.output_prefix : "";
    if (open_file && open_file->writer && open_file->prefix != prefix) open_file->finish();  // the caller moved on to another file
    if (!prefix.empty()) {
        if (open_file && open_file->writer) sh.writer = std::move(open_file->writer);  // continue the file an interrupt left open
        else sh.writer = std::make_unique<RecordWriter>(prefix, Game::name(), shape.bool_channels, shape.board, shape.scalar_count, shape.policy_len);
    }
    const uint64_t games_in_file = sh.writer ? sh.writer->game_count() : 0;  // max_games counts the file's games
    sh.job_count = size_t(std::max(1, c.gpu_batch / std::max(1, c.search_batch)));
    SlotTable<Game> local;
    SlotTable<Game>& per_thread = kept ? *kept : local;
    if (per_thread.empty()) {
        per_thread.resize(size_t(c.cpu_threads));
        for (int g = 0; g < games; g++)
            per_thread[size_t(g % c.cpu_threads)].push_back(std::make_unique<Slot<Game>>(c.seed * 1000003ull + uint64_t(g) * 7919ull + 1, size_t(c.cache_size), size_t(c.visits),
                                                                                         uint32_t(c.max_game_length)));
    } else {
        if (per_thread.size() != size_t(c.cpu_threads)) throw std::runtime_error("selfplay session: cpu_threads is a startup setting and cannot change between runs");
        games = 0;
        for (auto& slots : per_thread) {
            games += int(slots.size());
            // requests that were queued or never answered when the previous run stopped go back into the (new) queue; answered ones
            // are picked up by their generator thread as usual
            for (auto& sp : slots)
                if (sp->waiting && !sp->job.done.load(std::memory_order_acquire)) {
                    sh.queue.push_back(&sp->job);
                    sh.queued_positions += size_t(sp->job.n);
                }
        }
    }
    for (int t = 0; t < c.cpu_threads; t++) {
        sh.gen_mu.push_back(std::make_unique<std::mutex>());
        sh.gen_cv.push_back(std::make_unique<std::condition_variable>());
        sh.gen_signal.push_back(0);
    }
    for (auto& v : g_profile_cycles) v = 0;
    g_profile_cpu_s[0] = g_profile_cpu_s[1] = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int i = 0; i < c.gpu_threads; i++) threads.emplace_back([&, i] { executor_main(c.dummy_network ? nullptr : nets[size_t(i)].get(), sh, c, shape); });
    for (int t = 0; t < c.cpu_threads; t++) threads.emplace_back([&, t] { generator_main<Game>(t, per_thread[size_t(t)], sh, c); });
    bool interrupted = false;
    while (!sh.stop.load()) {
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (el >= c.duration_s || (c.max_moves > 0 && sh.moves.load() >= uint64_t(c.max_moves)) ||
            (c.max_games > 0 && games_in_file + sh.games.load() >= uint64_t(c.max_games)) || g_stop_requested.load())
            sh.stop.store(true);
        else if (open_file && g_interrupt_requested.load())
            interrupted = true, sh.stop.store(true);
    }
    sh.cv.notify_all();
    for (auto& cv : sh.gen_cv) cv->notify_all();
    for (auto& t : threads) t.join();
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out.games_written = sh.writer ? sh.writer->game_count() : 0;
    out.interrupted = 0;
    if (sh.writer) {
        // an interrupted run hands its unfinished file to the session: the next run (new network / settings) continues it
        if (interrupted && sh.error.empty() && (c.max_games <= 0 || sh.writer->game_count() < size_t(c.max_games))) {
            open_file->writer = std::move(sh.writer);
            open_file->prefix = prefix;
            out.interrupted = 1;
        } else {
            sh.writer->finish();
        }
    }
    if (std::getenv("KZB_SP_PROFILE")) {
        const double nodes = double(sh.real_evals.load() + sh.cached_evals.load());
        for (int i = 0; i < kSecCount; i++)
            std::fprintf(stderr, "[kzb selfplay] %-22s %8.1f cycles / node\n", kSectionNames[i], double(g_profile_cycles[i]) / std::max(nodes, 1.0));
        std::fprintf(stderr, "[kzb selfplay] thread CPU time: generators %.3f us / node, executors %.3f us / node\n",
                     1e6 * g_profile_cpu_s[0] / std::max(nodes, 1.0), 1e6 * g_profile_cpu_s[1] / std::max(nodes, 1.0));
    }
    if (!sh.error.empty()) throw std::runtime_error(sh.error);
    out.seconds = seconds;
    out.real_evals = sh.real_evals.load();
    out.cached_evals = sh.cached_evals.load();
    out.potential_evals = sh.potential_evals.load();
    out.batches = sh.batches.load();
    out.games_finished = sh.games.load();
    out.moves_played = sh.moves.load();
    out.root_visits = sh.root_visits.load();
    out.max_batch = sh.max_batch_seen.load();
    out.concurrent_games = uint64_t(games);
}

// deterministic stand-in network for the host-only search trace (the twin lives in oracle/mcts_oracle.py)
template <typename Game>
Eval pseudo_eval(const Game& b, int kind, std::vector<uint32_t>& scratch) {
    b.moves(scratch);
    const size_t n = scratch.size();
    Eval e;
    e.policy.resize(n);
    if (kind == 0) {  // DummyNetwork: uniform wdl and policy (rust/kz-core/src/network/dummy.rs:44-60)
        for (auto& p : e.policy) p = 1.0f / float(n);
        e.values = {0.0f, 1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f, 0.0f};
        return e;
    }
    const uint64_t h = b.hash();
    float sum = 0.0f;
    for (size_t i = 0; i < n; i++) {
        e.policy[i] = float((splitmix64(h + i + 1) >> 40) % 1000 + 1);
        sum += e.policy[i];
    }
    for (auto& p : e.policy) p /= sum;
    const float v = float(int((splitmix64(h ^ 0xABCDull) >> 40) % 2001) - 1000) / 1000.0f;
    e.values.value = v;
    e.values.win = (1.0f + v) * 0.5f * 0.8f;
    e.values.loss = (1.0f - v) * 0.5f * 0.8f;
    e.values.draw = 0.2f;
    e.values.moves_left = float((h >> 50) % 50);
    return e;
}

template <typename Game>
void trace_search(const kzb_selfplay_config& c, uint64_t game_seed, int plies, int eval_kind, kzb_mcts_trace_out& out) {
    const SearchSettings settings = search_settings(c);
    Game board = Game::start(game_seed);
    Rng rng(c.seed);
    std::vector<uint32_t> scratch;
    for (int i = 0; i < plies && !board.done(); i++) {  // walk into the game a little so that the root is not special
        board.moves(scratch);
        board.play(scratch[rng.gen_range(uint32_t(scratch.size()))]);
    }
    if (board.done()) throw std::runtime_error("trace: the game ended before the requested ply");
    Tree<Game> tree(board);
    tree.reserve(size_t(c.visits) * 48 + 64, size_t(c.visits) * 2 + 64);
    std::vector<Request<Game>> requests;
    uint64_t evals = 0;
    while (tree.root_visits() < uint64_t(c.visits)) {  // build_tree, generator_alphazero.rs:151-215 without the cache
        requests.clear();
        int terminal = 0;
        while (int(requests.size()) < c.search_batch && terminal < c.search_batch) {
            Request<Game> req;
            if (zero_step_gather(tree, settings, rng, req, scratch)) requests.push_back(std::move(req));
            else terminal++;
        }
        for (auto& req : requests) {
            Eval e = pseudo_eval(req.board, eval_kind, scratch);
            const float t = req.is_root() ? c.policy_temperature_root : c.policy_temperature_child;
            policy_softmax_temperature_in_place(e.policy.data(), e.policy.size(), t);
            zero_step_apply(tree, req.node, req.board.next_player(), e.values, e.policy.data(), e.policy.size());
            evals++;
        }
    }
    const int n_children = tree.root().child_count;
    if (n_children > out.capacity) throw std::runtime_error("trace: child capacity too small");
    out.n_children = n_children;
    std::vector<uint32_t> visits;
    tree.child_visits(Tree<Game>::kRoot, visits);
    for (int i = 0; i < n_children; i++) {
        out.child_visits[i] = visits[size_t(i)];
        out.child_moves[i] = tree.move_of(tree.root().child_start, n_children, i);
        out.child_policy[i] = tree.policy_of(tree.root().child_start)[i];
    }
    const ValuesPov v = pov(tree.root_values(), board.next_player());
    out.root_values[0] = v.value;
    out.root_values[1] = v.win;
    out.root_values[2] = v.draw;
    out.root_values[3] = v.loss;
    out.root_values[4] = v.moves_left;
    out.root_visits = tree.root_visits();
    out.tree_nodes = tree.size();
    out.evals = evals;
}

// a session = the games of one self-play server connection, kept between generations
struct SessionBase {
    virtual ~SessionBase() = default;
    virtual void run(int device, const void* onnx, size_t len, int precision, const kzb_selfplay_config& c, kzb_selfplay_stats& out) = 0;
    int game = -1;
};
template <typename Game>
struct Session : SessionBase {
    SlotTable<Game> slots;
    OpenRecordFile open_file;
    ~Session() override {
        try {
            open_file.finish();  // games already written are not lost when the connection ends in the middle of a file
        } catch (...) {
        }
    }
    void run(int device, const void* onnx, size_t len, int precision, const kzb_selfplay_config& c, kzb_selfplay_stats& out) override {
        run_selfplay<Game>(device, onnx, len, precision, c, out, &slots, &open_file);
    }
};

template <typename F>
int guarded(F&& f) {
    try {
        f();
        set_last_error("");
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    }
}

}  // namespace
}  // namespace selfplay
}  // namespace kzb

KZB_API void kzb_selfplay_default_config(kzb_selfplay_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    // python/main/loop_main_alpha.py:24-52 and UctWeights::default (node.rs:66-75)
    c->game = KZB_GAME_SYNTH_CHESS;
    c->visits = 800;
    c->search_batch = 16;
    c->gpu_batch = 1024;
    c->cpu_threads = 4;
    c->gpu_threads = 1;
    c->concurrent_games = 0;
    c->max_game_length = 400;
    c->cache_size = 800;
    c->zero_temp_move_count = 30;
    c->max_moves = 0;
    c->max_games = 0;
    c->part_iterations = 20;
    c->full_search_prob = 1.0f;
    c->duration_s = 5.0f;
    c->temperature = 1.0f;
    c->dirichlet_alpha = 0.03f;
    c->dirichlet_eps = 0.25f;
    c->policy_temperature_root = 1.4f;
    c->policy_temperature_child = 1.0f;
    c->exploration_weight = 2.0f;
    c->moves_left_weight = 0.03f;
    c->moves_left_clip = 20.0f;
    c->moves_left_sharpness = 0.5f;
    c->fpu_root = 0.1f;
    c->fpu_root_relative = 0;
    c->fpu_child = 0.0f;
    c->fpu_child_relative = 1;
    c->virtual_loss = 1.0f;
    c->q_mode_wdl = 1;
    c->draw_score = 0.0f;
    c->seed = 0;
}

KZB_API int kzb_selfplay_run(int device, const void* onnx_bytes, size_t onnx_len, int precision, const kzb_selfplay_config* config,
                             kzb_selfplay_stats* stats) {
    using namespace kzb::selfplay;
    return guarded([&] {
        if (!config || !stats) throw std::runtime_error("config and stats must not be NULL");
        if (!onnx_bytes && !config->dummy_network) throw std::runtime_error("onnx_bytes must not be NULL unless dummy_network is set");
        std::memset(stats, 0, sizeof(*stats));
        if (config->game == KZB_GAME_SYNTH_CHESS) run_selfplay<MaxMoves<SynthChess>>(device, onnx_bytes, onnx_len, precision, *config, *stats);
        else if (config->game == KZB_GAME_ATAXX7) run_selfplay<MaxMoves<Ataxx>>(device, onnx_bytes, onnx_len, precision, *config, *stats);
        else if (config->game == KZB_GAME_GO9) run_selfplay<MaxMoves<Go9>>(device, onnx_bytes, onnx_len, precision, *config, *stats);
        else if (config->game == KZB_GAME_CHESS) run_selfplay<MaxMoves<Chess>>(device, onnx_bytes, onnx_len, precision, *config, *stats);
        else if (config->game == KZB_GAME_GO9_TERRITORY) run_selfplay<MaxMoves<Go9Territory>>(device, onnx_bytes, onnx_len, precision, *config, *stats);
        else throw std::runtime_error("unknown game");
    });
}

KZB_API void kzb_selfplay_request_stop(void) { kzb::selfplay::g_stop_requested.store(true); }
KZB_API void kzb_selfplay_clear_stop(void) { kzb::selfplay::g_stop_requested.store(false); }
KZB_API void kzb_selfplay_request_interrupt(void) { kzb::selfplay::g_interrupt_requested.store(true); }
KZB_API void kzb_selfplay_clear_interrupt(void) { kzb::selfplay::g_interrupt_requested.store(false); }

struct kzb_selfplay_session {
    std::unique_ptr<kzb::selfplay::SessionBase> impl;
};

KZB_API int kzb_selfplay_session_create(int game, kzb_selfplay_session** out) {
    using namespace kzb::selfplay;
    return guarded([&] {
        if (!out) throw std::runtime_error("out must not be NULL");
        auto s = std::make_unique<kzb_selfplay_session>();
        if (game == KZB_GAME_SYNTH_CHESS) s->impl = std::make_unique<Session<MaxMoves<SynthChess>>>();
        else if (game == KZB_GAME_ATAXX7) s->impl = std::make_unique<Session<MaxMoves<Ataxx>>>();
        else if (game == KZB_GAME_GO9) s->impl = std::make_unique<Session<MaxMoves<Go9>>>();
        else if (game == KZB_GAME_CHESS) s->impl = std::make_unique<Session<MaxMoves<Chess>>>();
        else if (game == KZB_GAME_GO9_TERRITORY) s->impl = std::make_unique<Session<MaxMoves<Go9Territory>>>();
        else throw std::runtime_error("unknown game");
        s->impl->game = game;
        *out = s.release();
    });
}

KZB_API int kzb_selfplay_session_run(kzb_selfplay_session* session, int device, const void* onnx_bytes, size_t onnx_len, int precision,
                                     const kzb_selfplay_config* config, kzb_selfplay_stats* stats) {
    using namespace kzb::selfplay;
    return guarded([&] {
        if (!session || !config || !stats) throw std::runtime_error("session, config and stats must not be NULL");
        if (!onnx_bytes && !config->dummy_network) throw std::runtime_error("onnx_bytes must not be NULL unless dummy_network is set");
        if (config->game != session->impl->game) throw std::runtime_error("selfplay session: the game is a startup setting and cannot change between runs");
        std::memset(stats, 0, sizeof(*stats));
        session->impl->run(device, onnx_bytes, onnx_len, precision, *config, *stats);
    });
}

KZB_API void kzb_selfplay_session_destroy(kzb_selfplay_session* session) { delete session; }

KZB_API int kzb_mcts_trace(const kzb_selfplay_config* config, uint64_t game_seed, int plies, int eval_kind, kzb_mcts_trace_out* out) {
    using namespace kzb::selfplay;
    return guarded([&] {
        if (!config || !out || !out->child_visits || !out->child_moves || !out->child_policy) throw std::runtime_error("config / out must not be NULL");
        if (config->game == KZB_GAME_SYNTH_CHESS) trace_search<SynthChess>(*config, game_seed, plies, eval_kind, *out);
        else if (config->game == KZB_GAME_ATAXX7) trace_search<Ataxx>(*config, game_seed, plies, eval_kind, *out);
        else if (config->game == KZB_GAME_GO9) trace_search<Go9>(*config, game_seed, plies, eval_kind, *out);
        else if (config->game == KZB_GAME_CHESS) trace_search<Chess>(*config, game_seed, plies, eval_kind, *out);
        else if (config->game == KZB_GAME_GO9_TERRITORY) trace_search<Go9Territory>(*config, game_seed, plies, eval_kind, *out);
        else throw std::runtime_error("unknown game");
    });
}
