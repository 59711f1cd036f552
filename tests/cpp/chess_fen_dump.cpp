// For every FEN on stdin: the legal moves of kzero_b200/csrc/selfplay/chess_game.hpp as "mv <from> <to> <promotion> <flat index>"
// (absolute squares, a1 = 0; promotion q r b n or -), the en-passant plane and the scalars of the encoding, then "end".
// tests/test_chess_pairs.py holds the reference's known answers (rust/kz-core/tests/mapper/chess/pairs.rs) against this.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "../../kzero_b200/csrc/selfplay/chess_game.hpp"

int main() {
    using kzb::selfplay::Chess;
    std::string fen;
    while (std::getline(std::cin, fen)) {
        if (fen.empty()) continue;
        const Chess b = Chess::from_fen(fen);
        std::printf("fen %s\n", fen.c_str());
        b.legal_moves([&](const Chess::Mv& m) {
            std::printf("mv %d %d %c %u\n", int(m.from), int(m.to), m.promo ? "-pnbrqk"[m.promo] : '-', b.index_of(m));
            return true;
        });
        uint8_t bits[104];
        float sc[8];
        b.encode(bits, sc);
        unsigned long long planes[13];
        std::memcpy(planes, bits, 104);
        std::printf("ep_plane %llu\n", planes[12]);
        std::printf("own_pawns %llu\n", planes[0]);
        std::printf("scalars %g %g %g %g %g %g %g %g\n", sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], sc[6], sc[7]);
        std::printf("done %d\nend\n", int(b.done()));
    }
    return 0;
}
