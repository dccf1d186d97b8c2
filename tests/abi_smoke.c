/* abi_smoke.c -- a plain C99 caller of include/sabreur_gpu.h: proves the boundary is a C ABI (no C++ or torch types)
 * and shows the call sequence a host binds.  Built by tests/test_abi_cpu.py with gcc; run on a GPU by
 * tests/test_gpu_parity.py::test_plain_c_caller.  Usage: abi_smoke  -> prints "ok <n_records> <bin counts...>" */
#include <stdio.h>
#include <string.h>

#include "../include/sabreur_gpu.h"

int main(void) {
    static const char text[] =
        "@r1\nACGTAAAA\n+\nIIIIIIII\n"
        "@r2\nACGAAAAA\n+\nIIIIIIII\n"
        "@r3\nTTTTAAAA\n+\nIIIIIIII\n"
        "@r4\nGGGGNAAA\n+\nIIIIIIII\n";
    static const uint8_t barcodes[] = "ACGTGGGG"; /* two rows of 4 */
    sbr_config cfg;
    sbr_ctx* ctx = NULL;
    sbr_result res;
    uint8_t* buf = NULL;
    uint64_t cap = 0;
    uint32_t i;
    if (sbr_abi_version() != SBR_ABI_VERSION) return 2;
    if (sbr_device_count() <= 0) { fprintf(stderr, "no device: %s\n", sbr_last_error(NULL)); return 3; }
    memset(&cfg, 0, sizeof cfg);
    cfg.device = 0; cfg.fmt = SBR_FMT_AUTO; cfg.n_barcodes = 2; cfg.bc_len = 4; cfg.barcodes = barcodes;
    cfg.mismatch = 1; cfg.n_slots = 2; cfg.batch_bytes = 1u << 16;
    if (sbr_create(&ctx, &cfg) != SBR_OK) { fprintf(stderr, "create: %s\n", sbr_last_error(NULL)); return 4; }
    if (sbr_slot_buffer(ctx, 0, &buf, &cap) != SBR_OK || cap < sizeof text) return 5;
    memcpy(buf, text, sizeof text - 1);
    if (sbr_submit(ctx, 0, 0, sizeof text - 1, 0, 1) != SBR_OK) { fprintf(stderr, "submit: %s\n", sbr_last_error(ctx)); return 6; }
    if (sbr_wait(ctx, 0, &res) != SBR_OK) { fprintf(stderr, "wait: %s\n", sbr_last_error(ctx)); return 7; }
    printf("ok %u", res.n_records);
    for (i = 0; i < res.n_bins; ++i) printf(" %u", res.bin_count[i]);
    printf(" |");
    for (i = 0; i < res.n_records; ++i) printf(" %u", (unsigned)res.assign[i]);
    printf("\n");
    {   /* the device-resident entry points: the same batch from device memory, match + partition on the slot's own stream */
        sbr_ctx* ctx2 = NULL;
        void* d_text = NULL;
        uint64_t stats[2] = {0, 0};
        sbr_result res2;
        cfg.fmt = SBR_FMT_FASTQ;
        cfg.flags = SBR_CFG_OVERLAP;
        if (sbr_create(&ctx2, &cfg) != SBR_OK) { fprintf(stderr, "create 2: %s\n", sbr_last_error(NULL)); return 8; }
        if (sbr_dev_alloc(0, 4096, &d_text) != SBR_OK || sbr_dev_upload(0, d_text, buf, 4096) != SBR_OK) return 9;
        if (sbr_demux_device(ctx2, 0, (const uint8_t*)d_text, sizeof text - 1, 1, NULL) != SBR_OK) return 10;
        if (sbr_device_join(ctx2, NULL) != SBR_OK || sbr_scan_stats(ctx2, stats, 0) != SBR_OK) return 11;
        if (sbr_wait(ctx2, 0, &res2) != SBR_OK || res2.n_records != res.n_records) return 12;
        for (i = 0; i < res2.n_records; ++i)
            if (res2.assign[i] != res.assign[i]) return 13;
        if (stats[0] + stats[1] != 1) return 14;
        sbr_dev_free(0, d_text);
        sbr_destroy(ctx2);
    }
    sbr_destroy(ctx);
    return 0;
}
