/* Plain-C user of the drop-in boundary (include/svslam.h): the front-end seams of one stereo pair —
 * GFTT (src/frontend.cpp:51), LK left -> right (:105-109), triangulation (:174) — on a synthetic textured pair.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lstereovision-slam_b200 -lsvslam -Wl,-rpath,$PWD/stereovision-slam_b200 -lm -o demo
 *
 * Without a B200 svs_create() fails loudly (there is no CPU fallback) and the program exits with status 2. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "svslam.h"

#define W 620
#define H 188

static unsigned lcg(unsigned *s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

int main(void)
{
    svs_ctx *ctx = svs_create(0);
    if (!ctx) {
        fprintf(stderr, "svs_create failed: %s\n", svs_create_error());
        return 2;
    }
    /* blocky random texture; the right image is the left one shifted by a disparity of 6 px */
    unsigned char *left = (unsigned char *)malloc(W * H), *right = (unsigned char *)malloc(W * H);
    unsigned seed = 7;
    for (int by = 0; by < H; by += 6)
        for (int bx = 0; bx < W + 16; bx += 6) {
            unsigned char v = (unsigned char)(lcg(&seed) & 255);
            for (int y = by; y < by + 6 && y < H; y++)
                for (int x = bx; x < bx + 6; x++) {
                    if (x < W) left[y * W + x] = v;
                    if (x - 6 >= 0 && x - 6 < W) right[y * W + x - 6] = v;
                }
        }
    enum { MAXC = 150 };
    float xy[2 * MAXC], resp[MAXC], rxy[2 * MAXC];
    unsigned char status[MAXC], ok[MAXC];
    double xyz[3 * MAXC];
    int n = 0, rc;
    rc = svs_gftt_detect(ctx, left, W, H, W, NULL, 0, NULL, 0, MAXC, 0.01, 20.0, 32, xy, resp, &n);
    if (rc) { fprintf(stderr, "gftt: %s\n", svs_last_error(ctx)); return 1; }
    for (int i = 0; i < 2 * n; i++) rxy[i] = xy[i];                      /* initial flow = the left pixel (:98) */
    rc = svs_lk_track(ctx, left, right, W, H, W, xy, rxy, n, 11, 3, 30, 0.01, status);
    if (rc) { fprintf(stderr, "lk: %s\n", svs_last_error(ctx)); return 1; }
    const double K[4] = {359.428, 359.428, 303.5964, 92.60785};           /* KITTI seq-00, half resolution */
    rc = svs_triangulate(ctx, xy, rxy, n, K, K, 0.5371657, xyz, ok);
    if (rc) { fprintf(stderr, "triangulate: %s\n", svs_last_error(ctx)); return 1; }
    int good = 0;
    double zsum = 0;
    for (int i = 0; i < n; i++)
        if (status[i] && ok[i] && xyz[3 * i + 2] > 0) { good++; zsum += xyz[3 * i + 2]; }
    printf("%d corners, %d triangulated, mean depth %.2f m (expected %.2f m for a 6 px disparity), %lld kernels launched\n", n, good,
           good ? zsum / good : 0.0, K[0] * 0.5371657 / 6.0, svs_launch_count(ctx));
    free(left); free(right);
    svs_destroy(ctx);
    return good > 20 && fabs(zsum / good - K[0] * 0.5371657 / 6.0) < 1.0 ? 0 : 1;
}
