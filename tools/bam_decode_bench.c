/* Whole-file BAM decode through the C ABI from plain C, a fresh reader per pass (cold buffers: the real case), best of
 * five: open / decode / total including close.  CNVB_BAM_ZLIB=1, CNVB_BAM_SERIAL_WALK=1, CNVB_BAM_NO_THP=1 and
 * CNVB_BAM_CHUNK=<bytes> switch the decoder's parts for A/B runs.
 *   gcc -O2 -I include tools/bam_decode_bench.c -o /tmp/bam_decode_bench -L cnvetti_b200 -lcnvetti_b200 -Wl,-rpath,$PWD/cnvetti_b200
 *   /tmp/bam_decode_bench file.bam <threads>
 * A file to run it on: python -c "from cnvetti_b200 import synth; d = synth.synth_bam_records(22, 6686640, 1000000);
 *   synth.write_bam('/tmp/x.bam', [('chr22', 6686640)], [d], level=6, payload='random')"  */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <time.h>
#include "cnvetti_b200.h"
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
int main(int argc, char **argv) {
    int n = atoi(argv[2]);
    double best = 1e9, bopen = 0, bdec = 0;
    for (int it = 0; it < 5; ++it) {
        cnvb_bam *b;
        double t0 = now();
        if (cnvb_bam_open(argv[1], n, &b)) { printf("open failed\n"); return 1; }
        double t1 = now();
        if (cnvb_bam_decode(b, 0.6f, 0)) { printf("decode failed: %s\n", cnvb_bam_error(b)); return 1; }
        double t2 = now();
        cnvb_bam_close(b);
        double t3 = now();
        if (t3 - t0 < best) { best = t3 - t0; bopen = t1 - t0; bdec = t2 - t1; }
    }
    printf("threads %d cold: open %.1f ms decode %.1f ms total incl. close %.1f ms\n", n, bopen * 1e3, bdec * 1e3, best * 1e3);
    return 0;
}
