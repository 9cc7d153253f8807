/* A host without Python or CUDA headers running the DaNet network half through the C ABI (include/danet_b200.h):
 *
 *     gcc -O2 -Iinclude examples/net_host.c -o examples/net_host -Ldanet-densepose2smpl_b200 -ldanet_b200 \
 *         -Wl,-rpath,$PWD/danet-densepose2smpl_b200
 *     examples/net_host program.bin images.f32 para.f32 [graph]
 *
 * program.bin: written by DaNet.export_program(B, path); images.f32: raw fp32 [B,3,H,W] (normalised crops, what
 * demo.py:106 / eval.py:147 feed infer_net); para.f32: raw fp32 [B,229] = cam | shape | 24 rotation matrices. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "danet_b200.h"

static int fail(const char* what) {
    fprintf(stderr, "net_host: %s: %s\n", what, danet_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s program.bin images.f32 para.f32 [graph]\n", argv[0]);
        return 2;
    }
    danet_net_t net = NULL;
    if (danet_net_load_file(argv[1], &net) != 0) return fail("load");
    int32_t B = 0, chw[3] = {0, 0, 0}, n_out = 0, n_steps = 0;
    if (danet_net_info(net, &B, chw, &n_out, &n_steps) != 0) return fail("info");
    size_t n_in = (size_t)B * chw[0] * chw[1] * chw[2];
    float* img = (float*)malloc(n_in * sizeof(float));
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(img, sizeof(float), n_in, f) != n_in) { fprintf(stderr, "net_host: cannot read %zu floats from %s\n", n_in, argv[2]); return 1; }
    fclose(f);
    int flags = (argc > 4 && strcmp(argv[4], "graph") == 0) ? DANET_NET_GRAPH : 0;
    for (int rep = 0; rep < 2; ++rep)                   /* twice: the second call replays the captured graph */
        if (danet_net_infer_host(net, img, flags) != 0) return fail("infer");
    uint64_t bytes = 0; int32_t dims[4];
    if (danet_net_output(net, "para", NULL, &bytes, dims, NULL) != 0) return fail("output");
    float* para = (float*)malloc(bytes);
    if (danet_net_read_output(net, "para", para, bytes) != 0) return fail("read_output");
    f = fopen(argv[3], "wb");
    if (!f || fwrite(para, 1, bytes, f) != bytes) { fprintf(stderr, "net_host: cannot write %s\n", argv[3]); return 1; }
    fclose(f);
    printf("net_host: batch %d, input %dx%dx%d, %d steps, %d outputs; para [%d,%d] cam of image 0 = %g %g %g\n", B, chw[0], chw[1],
           chw[2], n_steps, n_out, dims[0], dims[1], para[0], para[1], para[2]);
    free(para); free(img);
    danet_net_destroy(net);
    return 0;
}
