// Self-play record files (`<prefix>.bin/.off/.json`), wire-compatible with the reference's trainer ("next" row N2).
//
// Restates the format of BinaryOutput (rust/kz-selfplay/src/binary_output.rs:128-297, scalar list :321-372): per position
//   26 f32 scalars | packed bools (BitBuffer::storage) | f32 input scalars | u32 policy indices | f32 policy values
// -- the packed (bits, scalars) part is byte-identical to the evaluator's input record -- plus one u64 byte offset per
// position in `.off` (followed by one u64 start index per game) and the metadata JSON.  Read back by the reference's
// own Python loader (python/lib/data/file.py:15-130, python/lib/data/position.py:34-103), which is what
// tests/test_selfplay_records.py does.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "mcts.hpp"

namespace kzb {
namespace selfplay {

// simulation.rs:14-24, with the board already reduced to what the file needs
struct RecordedPosition {
    std::vector<uint8_t> bits;      // BitBuffer::storage()
    std::vector<float> scalars;     // input scalars
    std::vector<uint32_t> indices;  // collect_policy_indices, binary_output.rs:299-315
    int next_player = 0;
    bool is_full_search = true;
    uint32_t played_index = 0;
    uint64_t zero_visits = 0;
    ValuesPov zero_values, net_values;
    std::vector<float> zero_policy, net_policy;
};

struct RecordedGame {
    std::vector<RecordedPosition> positions;
    RecordedPosition final_position;  // bits / scalars / next_player of the final board only
    bool final_done = true;           // false: stopped by the length limit
    int outcome = 0;                  // +1 A, 0 draw, -1 B (Outcome::Draw when not done, binary_output.rs:141)
};

class RecordWriter {
public:
    RecordWriter(const std::string& prefix, const std::string& game, int bool_channels, int board, int scalar_count, int policy_len)
        : prefix_(prefix), game_(game), bool_channels_(bool_channels), board_(board), scalar_count_(scalar_count), policy_len_(policy_len) {
        bin_ = std::fopen((prefix + ".bin").c_str(), "wb");
        off_ = std::fopen((prefix + ".off").c_str(), "wb");
        if (!bin_ || !off_) throw std::runtime_error("cannot create " + prefix + ".bin/.off");
    }
    ~RecordWriter() {
        if (bin_) std::fclose(bin_);
        if (off_) std::fclose(off_);
    }
    size_t game_count() const { return game_count_; }

    void append(const RecordedGame& g) {  // binary_output.rs:128-208
        const size_t game_id = game_count_, game_length = g.positions.size();
        game_start_indices_.push_back(uint64_t(position_count_));
        game_count_ += 1;
        position_count_ += 1 + game_length;
        max_len_ = std::max(max_len_, int(game_length));
        min_len_ = std::min(min_len_, int(game_length));
        const int start_player = g.positions.empty() ? g.final_position.next_player : g.positions[0].next_player;
        const int root_outcome = start_player == 0 ? g.outcome : -g.outcome;
        root_wdl_[root_outcome > 0 ? 0 : (root_outcome == 0 ? 1 : 2)] += 1;
        hit_move_limit_ += g.final_done ? 0 : 1;
        for (size_t pi = 0; pi < game_length; pi++) {
            const RecordedPosition& p = g.positions[pi];
            float kdl = 0.0f;  // kz-util/src/math.rs:7-11
            for (size_t i = 0; i < p.zero_policy.size(); i++) kdl += p.zero_policy[i] * std::log(p.zero_policy[i] / p.net_policy[i]);
            const int pov_outcome = p.next_player == 0 ? g.outcome : -g.outcome;
            const float moves_left = float(game_length + 1 - pi);
            float s[26] = {float(game_id), float(pi), float(game_length), float(p.zero_visits), p.is_full_search ? 1.0f : 0.0f, 0.0f, 0.0f,
                           0.0f, float(p.zero_policy.size()), float(p.played_index), kdl,
                           float(pov_outcome), pov_outcome > 0 ? 1.0f : 0.0f, pov_outcome == 0 ? 1.0f : 0.0f, pov_outcome < 0 ? 1.0f : 0.0f, moves_left,
                           p.zero_values.value, p.zero_values.win, p.zero_values.draw, p.zero_values.loss, p.zero_values.moves_left,
                           p.net_values.value, p.net_values.win, p.net_values.draw, p.net_values.loss, p.net_values.moves_left};
            write_position(s, p, p.indices, p.zero_policy);
        }
        const float nan = std::numeric_limits<float>::quiet_NaN();
        const int pov_outcome = g.final_position.next_player == 0 ? g.outcome : -g.outcome;
        float s[26] = {float(game_id), float(game_length), float(game_length), 0.0f, 0.0f, 1.0f, g.final_done ? 1.0f : 0.0f,
                       g.final_done ? 0.0f : 1.0f, 0.0f, -1.0f, nan,
                       float(pov_outcome), pov_outcome > 0 ? 1.0f : 0.0f, pov_outcome == 0 ? 1.0f : 0.0f, pov_outcome < 0 ? 1.0f : 0.0f, 0.0f,
                       nan, nan, nan, nan, nan, nan, nan, nan, nan, nan};
        write_position(s, g.final_position, {}, {});
    }

    void finish() {  // binary_output.rs:250-289
        if (finished_) throw std::logic_error("This output is already finished");
        finished_ = true;
        std::fwrite(game_start_indices_.data(), 8, game_start_indices_.size(), off_);
        std::fclose(bin_);
        std::fclose(off_);
        bin_ = off_ = nullptr;
        const std::string tmp = prefix_ + ".json.tmp";
        std::FILE* j = std::fopen(tmp.c_str(), "w");
        if (!j) throw std::runtime_error("cannot create " + tmp);
        const double gc = double(game_count_ ? game_count_ : 1);
        std::fprintf(j,
                     "{\n  \"game\": \"%s\",\n  \"input_bool_shape\": [%d, %d, %d],\n  \"input_scalar_count\": %d,\n  \"policy_shape\": [%d],\n"
                     "  \"game_count\": %zu,\n  \"position_count\": %zu,\n  \"includes_terminal_positions\": true,\n"
                     "  \"includes_game_start_indices\": true,\n  \"max_game_length\": %d,\n  \"min_game_length\": %d,\n"
                     "  \"root_wdl\": [%.9g, %.9g, %.9g],\n  \"hit_move_limit\": %.9g,\n  \"scalar_names\": [",
                     game_.c_str(), bool_channels_, board_, board_, scalar_count_, policy_len_, game_count_, position_count_,
                     game_count_ ? max_len_ : -1, game_count_ ? min_len_ : -1, root_wdl_[0] / gc, root_wdl_[1] / gc, root_wdl_[2] / gc,
                     hit_move_limit_ / gc);
        static const char* names[26] = {"game_id", "pos_index", "game_length", "zero_visits", "is_full_search", "is_final_position",
                                        "is_terminal", "hit_move_limit", "available_mv_count", "played_mv", "kdl_policy", "final_v",
                                        "final_wdl_w", "final_wdl_d", "final_wdl_l", "final_moves_left", "zero_v", "zero_wdl_w",
                                        "zero_wdl_d", "zero_wdl_l", "zero_moves_left", "net_v", "net_wdl_w", "net_wdl_d", "net_wdl_l",
                                        "net_moves_left"};
        for (int i = 0; i < 26; i++) std::fprintf(j, "%s\"%s\"", i ? ", " : "", names[i]);
        std::fprintf(j, "]\n}\n");
        std::fclose(j);
        if (std::rename(tmp.c_str(), (prefix_ + ".json").c_str()) != 0) throw std::runtime_error("cannot rename " + tmp);
    }

private:
    void write_position(const float* scalars26, const RecordedPosition& p, const std::vector<uint32_t>& indices, const std::vector<float>& values) {
        // binary_output.rs:210-247
        if (indices.size() != values.size()) throw std::logic_error("policy indices / values length mismatch");
        std::fwrite(&next_offset_, 8, 1, off_);
        auto put = [&](const void* d, size_t bytes) {
            if (bytes && std::fwrite(d, 1, bytes, bin_) != bytes) throw std::runtime_error("short write to " + prefix_ + ".bin");
            next_offset_ += bytes;
        };
        put(scalars26, 26 * 4);
        put(p.bits.data(), p.bits.size());
        put(p.scalars.data(), p.scalars.size() * 4);
        put(indices.data(), indices.size() * 4);
        put(values.data(), values.size() * 4);
    }

    std::string prefix_, game_;
    int bool_channels_, board_, scalar_count_, policy_len_;
    std::FILE *bin_ = nullptr, *off_ = nullptr;
    size_t game_count_ = 0, position_count_ = 0;
    int max_len_ = -1, min_len_ = std::numeric_limits<int>::max();
    double root_wdl_[3] = {0, 0, 0};
    uint64_t hit_move_limit_ = 0, next_offset_ = 0;
    std::vector<uint64_t> game_start_indices_;
    bool finished_ = false;
};

}  // namespace selfplay
}  // namespace kzb
