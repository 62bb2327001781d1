/*
 * bench_gpu.c -- C harness over libfse_b200.so mirroring the reference's block benchmark
 * (programs/bench.c: BMK_benchFiles :477-586, BMK_benchMem :301-474): same block split, same
 * parameters (maxSymbolValue 255, tableLog 12, slot = FSE_compressBound(chunk)), same 0 / 1 return
 * handling, a checksum self-check of the regenerated data, and the same result line -- except that
 * the two per-chunk loops (bench.c:353-364 and :389-424) are ONE batched call each
 * (FSEB200_compress_host / FSEB200_decompress_host).  Plain C: it sees only include/fse_b200.h.
 *
 *   build:  gcc -O2 -Iinclude programs/bench_gpu.c -Lfinitestateentropy_b200 -lfse_b200 \
 *               -Wl,-rpath,'$ORIGIN/../finitestateentropy_b200' -o programs/bench_gpu
 *   usage:  bench_gpu [-e|-h] [-i#] [-B#] FILE...     (-e FSE (default), -h Huff0, -i iterations, -B block size)
 *
 * Input files are read as the reference does (whole file, <= 1 GiB).  `probagen`-style inputs can be
 * produced with the reference's own programs/probagen or with tests/helpers.probagen().
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "fse_b200.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* FNV-1a 64: a self-check of the regenerated buffer (the reference uses XXH32, bench.c:311,444; any
 * checksum serves the purpose "dest == orig") */
static unsigned long long fnv64(const unsigned char* p, size_t n)
{
    unsigned long long h = 1469598103934665603ULL;
    size_t i;
    for (i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

int main(int argc, char** argv)
{
    int codec = 0, iterations = 4, a;
    size_t chunk = 32 * 1024;                               /* bench.c:98 */
    unsigned const nbSymbols = 255, tableLog = 12;          /* bench.c:113,569 */
    int nFiles = 0;
    if (FSEB200_device_count() < 1) { fprintf(stderr, "bench_gpu: no CUDA device (there is no CPU fallback)\n"); return 2; }
    for (a = 1; a < argc; a++) {
        if (argv[a][0] == '-') {
            if (argv[a][1] == 'e') codec = 0;
            else if (argv[a][1] == 'h') codec = 1;
            else if (argv[a][1] == 'i') iterations = atoi(argv[a] + 2);
            else if (argv[a][1] == 'B') chunk = (size_t)atoi(argv[a] + 2);
            continue;
        }
        {   const char* name = argv[a];
            FILE* f = fopen(name, "rb");
            size_t size, nb, slot, b, cSum = 0;
            unsigned char *orig, *cbuf, *dest;
            size_t *cSizes, *results;
            double bestC = 1e30, bestD = 1e30;
            int it, ok = 1;
            if (!f) { perror(name); return 1; }
            fseek(f, 0, SEEK_END); size = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
            if (size > ((size_t)1 << 30)) size = (size_t)1 << 30;
            nb = (size + chunk - 1) / chunk;
            slot = FSE_compressBound(chunk);
            orig = (unsigned char*)malloc(size + 64); dest = (unsigned char*)malloc(size + 64);
            cbuf = (unsigned char*)malloc(nb * slot + 64);
            cSizes = (size_t*)malloc(nb * sizeof(size_t)); results = (size_t*)malloc(nb * sizeof(size_t));
            if (!orig || !dest || !cbuf || !cSizes || !results) { fprintf(stderr, "not enough memory\n"); return 1; }
            if (fread(orig, 1, size, f) != size) { fprintf(stderr, "read error\n"); return 1; }
            fclose(f);
            for (it = 0; it < iterations; it++) {
                double t0 = now_s(), t1, t2;
                size_t r = FSEB200_compress_host(codec, cbuf, slot, cSizes, orig, size, chunk, nbSymbols, tableLog);
                t1 = now_s();
                if (FSE_isError(r)) { fprintf(stderr, "compress failed: %s\n", FSE_getErrorName(r)); return 1; }
                if (t1 - t0 < bestC) bestC = t1 - t0;
                memset(dest, 0, size);                      /* zeroing area, for checksum checking (bench.c:384) */
                t1 = now_s();
                r = FSEB200_decompress_host(codec, dest, size, chunk, cbuf, slot, cSizes, results, orig);
                t2 = now_s();
                if (FSE_isError(r)) { fprintf(stderr, "decompress failed: %s\n", FSE_getErrorName(r)); return 1; }
                if (t2 - t1 < bestD) bestD = t2 - t1;
                cSum = 0;
                for (b = 0; b < nb; b++) {
                    size_t const n = (b + 1) * chunk <= size ? chunk : size - b * chunk;
                    if (FSE_isError(cSizes[b])) { fprintf(stderr, "!!! Error compressing block %u !!!! => %s\n", (unsigned)b, FSE_getErrorName(cSizes[b])); return 1; }
                    cSum += cSizes[b] ? cSizes[b] : n;          /* bench.c:372-374 */
                    if (results[b] != n) { fprintf(stderr, "!! Error decompressing block %u of cSize %u !! => (%s)\n", (unsigned)b, (unsigned)cSizes[b], FSE_getErrorName(results[b])); ok = 0; break; }
                }
                if (fnv64(dest, size) != fnv64(orig, size)) { fprintf(stderr, "\n!!! %15s : Invalid Checksum !!!\n", name); ok = 0; }
                if (!ok) break;
            }
            if (ok)
                printf("%-17.17s : %9u -> %9u (%5.2f%%),%7.1f MB/s ,%7.1f MB/s   [%s, host buffers incl. PCIe]\n", name, (unsigned)size, (unsigned)cSum,
                       (double)cSum / (double)size * 100., (double)size / (1 << 20) / bestC, (double)size / (1 << 20) / bestD,
                       codec ? "Huff0" : "FSE");
            free(orig); free(dest); free(cbuf); free(cSizes); free(results);
            nFiles++;
        }
    }
    if (!nFiles) { fprintf(stderr, "usage: %s [-e|-h] [-i#] [-B#] FILE...\n", argv[0]); return 1; }
    return 0;
}
