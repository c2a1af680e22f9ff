/* dpi_caller.c -- stands in for the SystemVerilog side of tools/dpi/tb_dpi.sv (no simulator in this image): calls the
 * DPI-C shim exactly as the imported functions would be called and prints the testbench's report. */
#include <stdio.h>
#include <stdlib.h>
int rfb_dpi_open(const char *coe_path, long long size_range);
void rfb_dpi_close(void);
const char *rfb_dpi_error(void);
int rfb_dpi_scan2(const unsigned char *lo, const unsigned char *hi, int trace_entries, int capacity, unsigned int *n_records,
                  unsigned int *rec_stream, unsigned int *rec_pos, unsigned int *rec_state);
int rfb_dpi_cycles(const unsigned char *lo, const unsigned char *hi, int trace_entries, unsigned long long *cycles);

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    const int size_range = atoi(argv[2]), M = atoi(argv[5]), CAP = 1 << 20;
    unsigned char *lo = malloc((size_t)M + 1), *hi = malloc((size_t)M + 1);
    FILE *f = fopen(argv[3], "rb"); if (!f || fread(lo, 1, (size_t)M, f) != (size_t)M) return 1; fclose(f);
    f = fopen(argv[4], "rb"); if (!f || fread(hi, 1, (size_t)M, f) != (size_t)M) return 1; fclose(f);
    if (argc > 6) { puts("host only"); return 0; }
    if (rfb_dpi_open(argv[1], size_range)) { fprintf(stderr, "%s\n", rfb_dpi_error()); return 1; }
    unsigned n, *st = malloc(4u * CAP), *pos = malloc(4u * CAP), *state = malloc(4u * CAP);
    if (rfb_dpi_scan2(lo, hi, M, CAP, &n, st, pos, state)) { fprintf(stderr, "%s\n", rfb_dpi_error()); return 1; }
    unsigned *mc = calloc((size_t)size_range, 4), *mc2 = calloc((size_t)size_range, 4);
    for (unsigned k = 0; k < n && k < (unsigned)CAP; k++) { if (st[k] == 0) mc[state[k]]++; else mc2[state[k]]++; }
    for (int p = size_range - 1; p >= 0; p--) if (mc[p] & 0x3FF) printf("match_count[%d] = %u\n", p, mc[p] & 0x3FF);
    for (int p = size_range - 1; p >= 0; p--) if (mc2[p] & 0x3FF) printf("match_count_2[%d] = %u\n", p, mc2[p] & 0x3FF);
    unsigned long long cycles = 0;
    if (rfb_dpi_cycles(lo, hi, M, &cycles)) { fprintf(stderr, "%s\n", rfb_dpi_error()); return 1; }
    printf("Total no. cycles: %llu\n", cycles);
    rfb_dpi_close();
    return 0;
}
