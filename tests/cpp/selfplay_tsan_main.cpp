// Host-only ThreadSanitizer harness for the self-play driver (kzero_b200/csrc/selfplay/selfplay.cpp): generator and
// executor threads with the pseudo-network, the record writer on, no CUDA linked (the evaluator class is stubbed).
// Compiled with -fsanitize=thread and run by tests/test_host_units.py.
#include "../../kzero_b200/csrc/selfplay/selfplay.cpp"

namespace kzb {
Net::Net(int, const void*, size_t, int, int) { throw std::runtime_error("no GPU in this harness"); }
Net::~Net() {}
DeviceBuffer::~DeviceBuffer() {}
PinnedBuffer::~PinnedBuffer() {}
void Net::bind_mapper(int, int, int, int, int) {}
void Net::eval_packed(const uint8_t*, const float*, int, const uint32_t*, const uint32_t*, float*, float*, const uint8_t*) {}
void set_last_error(const std::string&) {}
}  // namespace kzb

#include <cstdio>

int main(int argc, char** argv) {
    kzb_selfplay_config c;
    kzb_selfplay_default_config(&c);
    c.visits = 100;
    c.search_batch = 8;
    c.gpu_batch = 128;
    c.cpu_threads = 3;
    c.gpu_threads = 2;
    c.duration_s = 2.0f;
    c.dummy_network = 2;
    c.executor_blocking_sync = 1;
    c.max_game_length = 12;
    c.output_prefix = argc > 1 ? argv[1] : "";
    kzb_selfplay_stats st;
    const int rc = kzb_selfplay_run(0, nullptr, 0, 1, &c, &st);
    std::printf("rc %d moves %llu evals %llu batches %llu games %llu\n", rc, (unsigned long long)st.moves_played,
                (unsigned long long)st.real_evals, (unsigned long long)st.batches, (unsigned long long)st.games_written);
    return rc != 0 || st.moves_played == 0 || st.games_written == 0;
}
