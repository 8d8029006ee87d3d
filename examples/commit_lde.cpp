// Host-side use of the C++ mirror: the calls the reference makes through ministark —
//   let lde  = trace.interpolate(domain).evaluate(lde_domain);        (Matrix)
//   let tree = MerkleTree::from_matrix(&lde);  tree.root();  tree.prove_rows(&positions)
// Build:  g++ -std=c++17 -Iinclude examples/commit_lde.cpp -Lsandstorm_b200 -lsandstorm_b200 -Wl,-rpath,$PWD/sandstorm_b200 -o commit_lde
#include <cstdio>

#include "sandstorm_b200.hpp"

using namespace sandstorm_b200;

int main() {
    try {
        Context ctx(0);
        const int n_cols = 3, log_n = 10;
        std::vector<std::vector<Felt>> cols(n_cols, std::vector<Felt>(size_t(1) << log_n));
        uint64_t s = 0x9E3779B97F4A7C15ull;                     // any residues below p = 2^251 + ... are valid elements
        for (auto &c : cols)
            for (auto &e : c) {
                for (int k = 0; k < 4; ++k) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; e[k] = s; }
                e[3] &= (uint64_t(1) << 58) - 1;
            }
        const Matrix trace = Matrix::from_columns(ctx, cols);
        const Matrix lde = trace.lde(1);
        const Matrix two_step = trace.interpolate().evaluate(1);
        const bool same = lde.column_to_host(0) == two_step.column_to_host(0);
        const MatrixMerkleTree tree = MatrixMerkleTree::from_matrix(lde, SS_TREE_KECCAK_M20);
        const Digest root = tree.root();
        const auto opened = tree.prove_rows({3, 1, 7});
        std::printf("lde == interpolate+evaluate: %s\nroot: ", same ? "yes" : "NO");
        for (uint8_t b : root) std::printf("%02x", b);
        std::printf("\nopened %zu elements, %zu path nodes\n", opened.first.size(), opened.second.size());
        return same ? 0 : 1;
    } catch (const Error &e) {
        std::fprintf(stderr, "%s (status %d)\n", e.what(), (int)e.status);
        return 2;
    }
}
