/* Plain-C caller of libmobicuda.so making the call sequence of the C# shim (csharp/MobiclipDecoder.cs) and of the
 * reference's front-ends (MobiConverter/Program.cs:64-73, 243-252):
 *     d = new MobiclipDecoder(W, H, Version);  per frame: d.Data = ...; d.Offset = 0; bmp = d.DecodeFrame();
 * Input: a file of frames, each preceded by its byte count (u32 LE).  Output (stdout): one line per frame,
 *     "<frame> <status> <offset_after> <quantizer> <fnv1a of Y|UV strided planes> <fnv1a of the BGRA bitmap>".
 * Built by tests/test_abi.py with gcc -std=c99 against include/mobicuda.h only: the boundary is a C ABI, no C++ / torch types. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mobicuda.h"

static unsigned long long fnv1a(const unsigned char* p, size_t n, unsigned long long h) {
    size_t i;
    for (i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

int main(int argc, char** argv) {
    unsigned w, h;
    int version, stride = 0, frame = 0;
    mobi_t* d = NULL;
    FILE* f;
    unsigned char *y, *uv, *bgra, *data = NULL;
    if (argc != 5) { fprintf(stderr, "usage: c_driver W H version frames.bin\n"); return 2; }
    w = (unsigned)atoi(argv[1]); h = (unsigned)atoi(argv[2]); version = atoi(argv[3]);
    if (mobi_create(w, h, version, 0, &d) != MOBI_OK) { fprintf(stderr, "mobi_create failed\n"); return 1; }
    mobi_get_state(d, NULL, NULL, &stride);
    y = malloc((size_t)stride * h); uv = malloc((size_t)stride * h / 2); bgra = malloc((size_t)w * h * 4);
    f = fopen(argv[4], "rb");
    if (!f || !y || !uv || !bgra) return 2;
    for (;;) {
        unsigned char hdr[4];
        unsigned len, q = 0;
        int off = 0, rc;
        unsigned long long hp = 14695981039346656037ULL, hb = 14695981039346656037ULL;
        if (fread(hdr, 1, 4, f) != 4) break;
        len = hdr[0] | hdr[1] << 8 | hdr[2] << 16 | (unsigned)hdr[3] << 24;
        data = realloc(data, len ? len : 1);
        if (fread(data, 1, len, f) != len) return 2;
        rc = mobi_decode_frame(d, data, (int)len, &off);              /* Data / Offset / DecodeFrame() */
        if (rc == MOBI_OK) {
            if (mobi_read_planes_strided(d, y, uv) != MOBI_OK) return 1;   /* Y[0], UV[0] */
            if (mobi_read_bgra(d, bgra, (int)w * 4) != MOBI_OK) return 1;  /* the returned Bitmap (Scan0, Stride) */
            mobi_get_state(d, &q, NULL, NULL);
            hp = fnv1a(uv, (size_t)stride * h / 2, fnv1a(y, (size_t)stride * h, hp));
            hb = fnv1a(bgra, (size_t)w * h * 4, hb);
        } else {
            fprintf(stderr, "frame %d: %s\n", frame, mobi_last_error(d));
        }
        printf("%d %d %d %u %016llx %016llx\n", frame, rc, off, q, hp, hb);
        frame++;
    }
    fclose(f);
    mobi_destroy(d);
    free(y); free(uv); free(bgra); free(data);
    return 0;
}
