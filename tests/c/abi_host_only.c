/* A plain C99 client of include/b2j.h: proves the header is valid C (what cgo / JNI / N-API stubs
 * compile against) and that the host-only entry points work without a GPU. */
#include <stdio.h>
#include <string.h>

#include "b2j.h"

int main(void)
{
    b2j_image_desc d;
    b2j_geometry g;
    uint8_t qy[64], qc[64];
    char msg[256];
    memset(&d, 0, sizeof d);
    d.width = 1920; d.height = 1080; d.subsampling = B2J_SUB_420; d.n_frames = 256;
    if (b2j_get_geometry(&d, &g) != B2J_OK) return 1;
    if (g.mcus_x != 120 || g.mcus_y != 68 || g.blocks_per_mcu != 6 || g.blocks_per_frame != 48960 || g.header_bytes != 623) return 2;
    if (b2j_quant_tables(85, qy, qc) != B2J_OK || qy[0] != 5 || qc[0] != 5) return 3;      /* the model's tables */
    if (b2j_quant_tables(0, qy, qc) == B2J_OK) return 4;
    if (b2j_last_error(msg, sizeof msg) == 0) return 5;
    if (b2j_default_out_cap(&d) < 623 + 2) return 6;
    if (b2j_entropy_workspace_bytes_opts(&d, b2j_default_out_cap(&d), B2J_ZRL_EAGER | B2J_OPT_RESTART(120)) <=
        b2j_entropy_workspace_bytes(&d, b2j_default_out_cap(&d))) return 7;
    d.width = 70000;
    if (b2j_get_geometry(&d, &g) == B2J_OK) return 8;
    printf("%s geometry %dx%d MCUs, %lld blocks per frame\n", b2j_version(), g.mcus_x, g.mcus_y, (long long)g.blocks_per_frame);
    return 0;
}
