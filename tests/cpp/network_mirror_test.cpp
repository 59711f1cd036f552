// Drives include/kzb200.hpp -- the C++ mirror of kz-core's `Network` interface -- the way the reference's executor drives `CudaNetwork`:
// boards in, one ZeroEvaluation per board out.  The boards are positions of the bundled games (random playouts), the mapper adapts a
// game to the BoardMapper concept.  Prints, per board, the packed record it sent (so the Python test can send the SAME record through
// the Python mirror and compare bit for bit) and the evaluation it got; then exercises the error behaviour (cudnn.rs:58,
// common.rs:165-198).  Compiled and run by tests/test_cpp_mirror.py:
//     network_mirror_test <net.onnx> <chess|go-9|ataxx-7> <boards> <seed> <max_batch>
// Exit code 0: evaluated; 3: the library reported an error while the network was created (message on stdout) -- what happens on a
// machine without a CUDA device, where the product must fail loudly.
#include <array>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/kzb200.hpp"
#include "../../kzero_b200/csrc/selfplay/chess_game.hpp"
#include "../../kzero_b200/csrc/selfplay/games.hpp"
#include "../../kzero_b200/csrc/selfplay/mcts.hpp"

using namespace kzb::selfplay;

// a bundled game as BoardMapper<Game> (rust/kz-core/src/mapping/mod.rs:9-36)
template <typename Game>
struct GameMapper {
    std::array<int, 3> input_bool_shape() const {
        const GameShape s = Game::shape();
        return {s.bool_channels, s.board, s.board};
    }
    int input_scalar_count() const { return Game::shape().scalar_count; }
    int policy_len() const { return Game::shape().policy_len; }
    void encode_input(uint8_t* bits, float* scalars, const Game& board) const { board.encode(bits, scalars); }
    void available_move_indices(const Game& board, std::vector<uint32_t>& out) const {
        out.clear();
        if (board.done()) return;  // available_moves() of a finished board is an error: empty policy (common.rs:77)
        board.moves(out);
        for (auto& mv : out) mv = board.move_to_index(mv);
    }
};

template <typename Game>
int run(const std::vector<char>& onnx, int n, uint64_t seed, int max_batch) {
    using Net = kzb200::B200Network<Game, GameMapper<Game>>;
    std::vector<Game> boards;
    Rng rng(seed);
    std::vector<uint32_t> moves;
    for (int i = 0; i < n; i++) {  // positions at different depths of random games; the last one played to its end when it ends early
        Game b = Game::start(seed + uint64_t(i));
        const int plies = int(rng.gen_range(60));
        for (int p = 0; p < plies && !b.done(); p++) {
            b.moves(moves);
            b.play(moves[rng.gen_range(uint32_t(moves.size()))]);
        }
        boards.push_back(b);
    }
    try {
        Net net(GameMapper<Game>(), onnx.data(), onnx.size(), max_batch, 0);
        std::printf("max_batch_size %d\n", net.max_batch_size());
        const std::vector<kzb200::ZeroEvaluation> evals = net.evaluate_batch(boards);
        const GameShape shape = Game::shape();
        std::vector<uint8_t> bits(size_t(shape.bits_bytes()));
        std::vector<float> scalars(size_t(shape.scalar_count));
        for (size_t i = 0; i < boards.size(); i++) {
            std::fill(bits.begin(), bits.end(), 0);
            net.mapper().encode_input(bits.data(), scalars.data(), boards[i]);
            net.mapper().available_move_indices(boards[i], moves);
            std::printf("board %zu done %d\nbits", i, int(boards[i].done()));
            for (uint8_t v : bits) std::printf(" %u", unsigned(v));
            std::printf("\nscalars");
            for (float v : scalars) std::printf(" %.9g", v);
            std::printf("\nindices");
            for (uint32_t v : moves) std::printf(" %u", v);
            const kzb200::ZeroEvaluation& e = evals[i];
            std::printf("\nvalues %.9g %.9g %.9g %.9g %.9g\npolicy", e.values.value, e.values.wdl.win, e.values.wdl.draw, e.values.wdl.loss, e.values.moves_left);
            for (float v : e.policy) std::printf(" %.9g", v);
            std::printf("\n");
        }
        // the single-board form answers like its row of the batch (rows are independent of the batch, bit for bit)
        const kzb200::ZeroEvaluation one = net.evaluate(boards[0]);
        std::printf("single_equals_row %d\n", int(one.policy == evals[0].policy && one.values.value == evals[0].values.value));
        // cudnn.rs:58: more boards than max_batch_size
        try {
            std::vector<Game> too_many(size_t(max_batch) + 1, boards[0]);
            net.evaluate_batch(too_many);
            std::printf("too_many no error\n");
        } catch (const kzb200::Error& e) {
            std::printf("too_many error %s\n", e.what());
        }
        // an empty batch is an empty answer
        std::printf("empty %zu\n", net.evaluate_batch(std::vector<Game>()).size());
    } catch (const kzb200::Error& e) {
        std::printf("error %s\n", e.what());
        return 3;
    }
    // check_graph_shapes (common.rs:171-174): a mapper of another game does not fit this graph
    try {
        using Other = typename std::conditional<std::is_same<Game, Ataxx>::value, Chess, Ataxx>::type;
        kzb200::B200Network<Other, GameMapper<Other>> wrong(GameMapper<Other>(), onnx.data(), onnx.size(), max_batch, 0);
        std::printf("mismatch no error\n");
    } catch (const kzb200::Error& e) {
        std::printf("mismatch error %s\n", e.what());
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 6) return std::printf("usage: network_mirror_test <net.onnx> <game> <boards> <seed> <max_batch>\n"), 2;
    std::ifstream f(argv[1], std::ios::binary);
    const std::vector<char> onnx((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const std::string game = argv[2];
    const int n = std::atoi(argv[3]), max_batch = std::atoi(argv[5]);
    const uint64_t seed = std::strtoull(argv[4], nullptr, 10);
    std::printf("devices %d\n", kzb200::device_count());
    if (game == "chess") return run<Chess>(onnx, n, seed, max_batch);
    if (game == "go-9") return run<Go9>(onnx, n, seed, max_batch);
    if (game == "ataxx-7") return run<Ataxx>(onnx, n, seed, max_batch);
    return std::printf("unknown game %s\n", game.c_str()), 2;
}
