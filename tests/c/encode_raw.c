/* Pure C client of the host-buffer encoder: raw interleaved RGB file -> JFIF file(s).
 *   encode_raw in.rgb width height n_frames subsampling(444|422|420) quality zrl(0|1) restart out_prefix
 * Uses only include/b2j.h + libb2j.so (no Python, no torch): what a cgo / JNI binding would call. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2j.h"

static int fail(const char *what, int rc)
{
    char msg[512];
    b2j_last_error(msg, sizeof msg);
    fprintf(stderr, "%s: rc %d: %s\n", what, rc, msg);
    return 1;
}

int main(int argc, char **argv)
{
    if (argc != 10) { fprintf(stderr, "usage: %s in.rgb width height n sub quality zrl restart out_prefix\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), n = atoi(argv[4]), sub = atoi(argv[5]), quality = atoi(argv[6]);
    const int opts = atoi(argv[7]) | B2J_OPT_RESTART(atoi(argv[8]));
    const size_t frame = (size_t)w * h * 3;
    uint8_t *rgb = (uint8_t *)malloc(frame * n);
    FILE *f = fopen(argv[1], "rb");
    if (!rgb || !f || fread(rgb, 1, frame * n, f) != frame * n) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);

    b2j_image_desc d;
    memset(&d, 0, sizeof d);
    d.width = w; d.height = h; d.subsampling = sub; d.n_frames = 1;
    const int64_t cap = b2j_default_out_cap(&d) * 4;
    b2j_encoder *enc = NULL;
    int rc = b2j_encoder_create(&enc, &d, quality, opts, cap, n < 4 ? n : 4);
    if (rc) return fail("b2j_encoder_create", rc);
    uint8_t *out = (uint8_t *)malloc((size_t)cap * n);
    int64_t *sizes = (int64_t *)malloc(sizeof(int64_t) * n);
    rc = b2j_encoder_encode_host(enc, rgb, n, out, sizes);          /* pageable buffers: staged through pinned slots */
    if (rc) return fail("b2j_encoder_encode_host", rc);
    b2j_encoder_destroy(enc);
    for (int i = 0; i < n; i++) {
        char name[1024];
        snprintf(name, sizeof name, "%s%d.jpg", argv[9], i);
        FILE *o = fopen(name, "wb");
        if (!o || fwrite(out + (size_t)i * cap, 1, (size_t)sizes[i], o) != (size_t)sizes[i]) { fprintf(stderr, "cannot write %s\n", name); return 2; }
        fclose(o);
    }
    printf("%d frame(s) encoded\n", n);
    free(rgb); free(out); free(sizes);
    return 0;
}
