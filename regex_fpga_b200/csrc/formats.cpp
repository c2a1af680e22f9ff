// formats.cpp -- the reference's two text formats.
#include "host.h"
#include "../../include/regex_fpga_b200.h"
#include <cstdio>
#include <cstring>

namespace rfb {

static bool read_all(const std::string &path, std::string &out) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    bool ok = !std::ferror(f);
    std::fclose(f);
    return ok;
}

static inline int hexdigit(unsigned char ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    ch |= 0x20;
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    return -1;
}
static inline bool is_sep(unsigned char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == ','; }

// Grammar accepted (Xilinx COE, radix 16 only):
//   memory_initialization_radix = 16 ;  memory_initialization_vector = W (sep W)* [;]
// with W exactly 32 hex digits = one 128-bit BRAM line.  The first 8 digits of W are
// rd_bus[127:96] = cache[0] (Design/FPGA.v:884), so they become entry 4*line + 0.
int coe_parse_file(const std::string &path, std::vector<uint32_t> &entries, std::string &err) {
    std::string txt;
    if (!read_all(path, txt)) { err = "cannot read " + path; return RFB_E_IO; }
    size_t rad = txt.find("memory_initialization_radix");
    if (rad != std::string::npos) {
        size_t eq = txt.find('=', rad);
        if (eq == std::string::npos || std::atoi(txt.c_str() + eq + 1) != 16) {
            err = path + ": only memory_initialization_radix=16 is supported";
            return RFB_E_FORMAT;
        }
    }
    size_t vec = txt.find("memory_initialization_vector");
    if (vec == std::string::npos) { err = path + ": no memory_initialization_vector"; return RFB_E_FORMAT; }
    size_t pos = txt.find('=', vec);
    if (pos == std::string::npos) { err = path + ": malformed vector keyword"; return RFB_E_FORMAT; }
    pos++;
    entries.clear();
    entries.reserve(txt.size() / 8);
    const size_t n = txt.size();
    while (pos < n) {
        while (pos < n && is_sep((unsigned char)txt[pos])) pos++;
        if (pos >= n || txt[pos] == ';') break;
        uint32_t w[4] = {0, 0, 0, 0};
        int digits = 0;
        while (pos < n) {
            int v = hexdigit((unsigned char)txt[pos]);
            if (v < 0) break;
            if (digits < 32) w[digits >> 3] = (w[digits >> 3] << 4) | (uint32_t)v;
            digits++; pos++;
        }
        if (digits != 32) {
            err = path + ": word " + std::to_string(entries.size() / 4) + " has " +
                  std::to_string(digits) + " hex digits (want 32)";
            return RFB_E_FORMAT;
        }
        if (pos < n && !is_sep((unsigned char)txt[pos]) && txt[pos] != ';') {
            err = path + ": unexpected character after word " + std::to_string(entries.size() / 4);
            return RFB_E_FORMAT;
        }
        entries.insert(entries.end(), w, w + 4);
    }
    if (entries.empty()) { err = path + ": empty vector"; return RFB_E_FORMAT; }
    return RFB_OK;
}

int coe_write_file(const std::string &path, const uint32_t *e, size_t n, int style, std::string &err) {
    if (n % 4 != 0) { err = "entry count must be a multiple of 4 (128-bit lines)"; return RFB_E_INVALID; }
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return RFB_E_IO; }
    std::fputs("memory_initialization_radix=16;\nmemory_initialization_vector=", f);
    for (size_t l = 0; l < n / 4; l++) {
        if (l) std::fputc(style == 1 ? ' ' : '\n', f);
        std::fprintf(f, "%08x%08x%08x%08x", e[4 * l], e[4 * l + 1], e[4 * l + 2], e[4 * l + 3]);
    }
    if (style == 1) std::fputs(";\n", f);
    bool ok = !std::ferror(f);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { err = "write error on " + path; return RFB_E_IO; }
    return RFB_OK;
}

// One token per symbol, 1-2 hex digits, whitespace separated, token k -> index k.
int mem_parse_file(const std::string &path, std::vector<uint8_t> &bytes, std::string &err) {
    std::string txt;
    if (!read_all(path, txt)) { err = "cannot read " + path; return RFB_E_IO; }
    bytes.clear();
    bytes.reserve(txt.size() / 2);
    size_t pos = 0, n = txt.size();
    while (pos < n) {
        while (pos < n && (is_sep((unsigned char)txt[pos]) && txt[pos] != ',')) pos++;
        if (pos >= n) break;
        int v = 0, digits = 0;
        while (pos < n) {
            int d = hexdigit((unsigned char)txt[pos]);
            if (d < 0) break;
            v = v * 16 + d; digits++; pos++;
        }
        bool ends_ok = pos >= n || txt[pos] == ' ' || txt[pos] == '\t' || txt[pos] == '\n' || txt[pos] == '\r';
        if (digits < 1 || digits > 2 || !ends_ok) {
            err = path + ": token " + std::to_string(bytes.size()) + " is not a 1-2 digit hex byte";
            return RFB_E_FORMAT;
        }
        bytes.push_back((uint8_t)v);
    }
    return RFB_OK;
}

int mem_write_file(const std::string &path, const uint8_t *b, size_t n, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return RFB_E_IO; }
    for (size_t k = 0; k < n; k++) std::fprintf(f, "%x\n", b[k]);  // no leading zero, as shipped
    bool ok = !std::ferror(f);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { err = "write error on " + path; return RFB_E_IO; }
    return RFB_OK;
}

// row_ptr[0..size] then row_ptr[size] transitions, zero-padded to a full line
// (Design/FPGA.v:773,782,793).  Unique `size` with: E[0]==0, prefix non-decreasing, 0..3 zero pad.
int64_t detect_size(const uint32_t *E, size_t n) {
    if (n < 2 || E[0] != 0) return -1;
    int64_t found = -1;
    for (size_t size = 1; size < n; size++) {
        if (E[size] < E[size - 1]) break;
        const uint64_t used = (uint64_t)size + 1 + E[size];
        if (used > n || n - used > 3) continue;
        bool zero_pad = true;
        for (size_t j = (size_t)used; j < n; j++) zero_pad = zero_pad && E[j] == 0;
        if (!zero_pad) continue;
        if (found >= 0) return -1;
        found = (int64_t)size;
    }
    return found;
}

}  // namespace rfb
