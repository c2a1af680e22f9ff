// fuzz_imagefile.cpp -- dev tool: mutate an execution-image file under a re-sealed checksum and feed it to
// plan_read() under AddressSanitizer.  Every mutant must be either refused or accepted without a memory error.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -I regex_fpga_b200/csrc tools/dev/fuzz_imagefile.cpp \
//       regex_fpga_b200/csrc/{imagefile,image,nfa,formats,parts}.cpp -o /tmp/fuzz_imagefile
//   /tmp/fuzz_imagefile file.rfbimg [n_mutants] [seed]
#include "host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
using namespace rfb;

static uint64_t fnv1a64(const uint8_t *p, size_t n) { uint64_t h = 0xcbf29ce484222325ull; for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; } return h; }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<uint8_t> base; { uint8_t buf[65536]; size_t k; while ((k = fread(buf, 1, sizeof buf, f)) > 0) base.insert(base.end(), buf, buf + k); } fclose(f);
    const int n = argc > 2 ? atoi(argv[2]) : 200;
    std::mt19937_64 rng(argc > 3 ? atoll(argv[3]) : 1);
    uint64_t n_entries; memcpy(&n_entries, &base[16], 8);
    const size_t tables = 32 + 4 * n_entries;
    int accepted = 0, refused = 0;
    for (int i = 0; i < n; i++) {
        std::vector<uint8_t> m = base;
        const int edits = 1 + (int)(rng() % 4);
        for (int e = 0; e < edits; e++) {
            // mostly the table area (the CSR is covered by the equivalence proof anyway), sometimes anywhere
            size_t at = (rng() % 8) ? tables + rng() % (m.size() - 8 - tables) : rng() % (m.size() - 8);
            switch (rng() % 4) {
                case 0: m[at] ^= (uint8_t)(1u << (rng() % 8)); break;
                case 1: m[at] = (uint8_t)rng(); break;
                case 2: m[at] = 0xFF; break;
                default: if (at + 4 < m.size() - 8) { uint32_t v = (uint32_t)rng(); memcpy(&m[at], &v, 4); } break;
            }
        }
        if (rng() % 16 == 0) m.resize(m.size() - (rng() % 64) * 4 - 8 + 8);   // occasional truncation
        const uint64_t sum = fnv1a64(m.data(), m.size() - 8);
        memcpy(&m[m.size() - 8], &sum, 8);
        const char *tmp = "/tmp/fuzz_mutant.rfbimg";
        FILE *g = fopen(tmp, "wb"); fwrite(m.data(), 1, m.size(), g); fclose(g);
        Plan plan; std::string err;
        if (plan_read(tmp, plan, err) == 0) accepted++; else refused++;
    }
    printf("mutants %d refused %d accepted %d\n", n, refused, accepted);
    return 0;
}
