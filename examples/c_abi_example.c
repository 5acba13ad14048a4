/* Plain-C use of libraytracegr_cuda (include/raytracegr_cuda.h): example2 of the reference
 * (src/RayTraceGR.jl:578-612) rendered through rtgr_render into an 8-bit RGB buffer and written as a PPM.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Lraytracegr.jl_b200/csrc -lraytracegr_cuda -o c_abi_example
 *   LD_LIBRARY_PATH=raytracegr.jl_b200/csrc ./c_abi_example out.ppm
 *
 * Without a Blackwell GPU rtgr_create fails (there is no CPU fallback); the program then prints the
 * library's message and exits with status 2 -- which is what tests/test_abi.py checks on the CPU box. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "raytracegr_cuda.h"

int main(int argc, char** argv) {
    rtgr_params p;
    rtgr_object objs[3];
    rtgr_camera cam;
    rtgr_stats st;
    rtgr_ctx* ctx = NULL;
    unsigned char* rgb;
    const int ni = 200, nj = 200;

    printf("libraytracegr_cuda version %d\n", rtgr_version());
    rtgr_default_params(&p, RTGR_KERR_SCHILD);                 /* M = 1, a = 0, tol = eps^(3/4), lambda in (0, 100) */

    memset(objs, 0, sizeof(objs));
    objs[0].kind = RTGR_SPHERE; objs[0].vel[0] = 1.0; objs[0].radius = -10.0;                 /* caelum  (src:582) */
    objs[1].kind = RTGR_PLANE;  objs[1].time = -20.0;                                         /* frustum (src:583) */
    objs[2].kind = RTGR_SPHERE; objs[2].pos[1] = 4.0; objs[2].vel[0] = 1.0; objs[2].radius = 0.5; /* sphere (src:584) */

    memset(&cam, 0, sizeof(cam));
    cam.pos[1] = 4.0; cam.pos[2] = -2.0;                       /* src:590 */
    cam.widthx[1] = 1.0; cam.widthy[3] = 1.0; cam.normal[2] = 1.0;
    cam.ni = ni; cam.nj = nj;

    if (rtgr_create(&ctx, NULL, 1) != 0) {
        fprintf(stderr, "rtgr_create: %s\n", rtgr_last_error());
        return 2;
    }
    rgb = (unsigned char*)malloc((size_t)ni * nj * 3);
    if (rtgr_render(ctx, &p, objs, 3, &cam, rgb, NULL, NULL, NULL, NULL, NULL, &st) != 0) {
        fprintf(stderr, "rtgr_render: %s\n", rtgr_last_error());
        rtgr_destroy(ctx);
        return 1;
    }
    printf("%llu rays, %llu RHS evaluations, kernel %.3f ms\n", (unsigned long long)st.rays,
           (unsigned long long)st.rhs_evals, st.kernel_ms);
    if (argc > 1) {
        FILE* f = fopen(argv[1], "wb");
        if (f) { fprintf(f, "P6\n%d %d\n255\n", ni, nj); fwrite(rgb, 3, (size_t)ni * nj, f); fclose(f); }
    }
    free(rgb);
    rtgr_destroy(ctx);
    return 0;
}
