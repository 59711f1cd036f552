// Prints the flat policy table of kzero_b200/csrc/selfplay/chess_game.hpp, one "from to promotion" line per index
// (promotion: q r b n or -), for tests/test_host_units.py to compare with the reference's table.
#include <cstdio>

#include "../../kzero_b200/csrc/selfplay/chess_game.hpp"

int main() {
    const auto& t = kzb::selfplay::chess_detail::flat_moves();
    for (int i = 0; i < 1880; i++) std::printf("%d %d %c\n", t.from[i], t.to[i], t.promo[i] ? "-pnbrqk"[t.promo[i]] : '-');
    return 0;
}
