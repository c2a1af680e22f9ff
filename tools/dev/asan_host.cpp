// asan_host.cpp -- dev tool: run the host-side model (COE parse, NFA validation, scan plan incl. start DFA and
// verifier, edge-grouped CSR, image file round trip) over .coe files under AddressSanitizer / UBSan.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -I regex_fpga_b200/csrc tools/dev/asan_host.cpp \
//       regex_fpga_b200/csrc/{imagefile,image,nfa,formats,parts,ecsr}.cpp -o /tmp/asan_host
#include "host.h"
#include <cstdio>
#include <cstdlib>
using namespace rfb;
int main(int argc, char **argv) {
    int ok = 0, not_ok = 0;
    for (int i = 1; i < argc; i++) {
        std::vector<uint32_t> E; std::string err;
        if (coe_parse_file(argv[i], E, err)) { printf("%s: parse: %s\n", argv[i], err.c_str()); continue; }
        for (uint32_t budget : {0u, 20u, 16384u}) {
            ImageOptions opt; opt.dfa_max_states = budget; opt.accel = budget ? 1 : 0; opt.bucket_bits = -1;
            opt.max_bytes = 227 * 1024 - 32 * 1024 - 64;
            Plan plan;
            if (plan_build(E.data(), E.size(), -1, opt, true, plan, err)) { printf("%s: build: %s\n", argv[i], err.c_str()); continue; }
            for (auto &p : plan.parts) { Ecsr ec; if (ecsr_build(p.sub, ec, err)) printf("%s: ecsr: %s\n", argv[i], err.c_str()); (p.img.ok ? ok : not_ok)++; }
            if (plan_write(plan, "/tmp/asan_plan.rfbimg", err)) { printf("write: %s\n", err.c_str()); continue; }
            Plan back;
            if (plan_read("/tmp/asan_plan.rfbimg", back, err)) printf("%s: read back: %s\n", argv[i], err.c_str());
        }
    }
    printf("parts with tables %d, without %d\n", ok, not_ok);
    return 0;
}
