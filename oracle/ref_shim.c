/*
 * ref_shim.c -- TEST/BENCH INFRASTRUCTURE ONLY (not product code).
 *
 * Thin driver compiled *together with the unmodified reference sources* (which stay where
 * they lie under /root/reference/lib; nothing is copied) into oracle/_ref/libfse_ref.so by
 * oracle/Makefile.  It only adds what the reference harness does around the library:
 * the per-chunk loops of programs/bench.c:353-364 (compress) and :389-424 (decompress),
 * run over disjoint block ranges by N pthreads so that bench.py can report an all-cores
 * CPU baseline.  Every reference entry point is re-entrant
 * (stack workspaces only), so threading over blocks is legitimate.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the resulting library.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fse.h"
#include "huf.h"
#include "fseU16.h"

enum { SHIM_FSE = 0, SHIM_HUF = 1, SHIM_FSEU16 = 2 };

typedef struct {
    int codec, decode;
    const unsigned char* src;     /* original data (also used for raw / RLE blocks on decode) */
    unsigned char* out;           /* decode: regenerated data */
    unsigned char* cbuf;          /* compressed slots */
    size_t total, blockSize, slot;
    size_t* csizes;               /* compress: written; decode: read */
    size_t* results;              /* decode: regenerated size or error code per block */
    unsigned msv, tlog;
    size_t first, last;           /* block range [first,last) */
} job_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    size_t b;
    for (b = j->first; b < j->last; b++) {
        size_t const off = b * j->blockSize;
        size_t const n = (off + j->blockSize <= j->total) ? j->blockSize : j->total - off;
        unsigned char* const c = j->cbuf + b * j->slot;
        if (!j->decode) {
            size_t r;
            if (j->codec == SHIM_FSE)      r = FSE_compress2(c, j->slot, j->src + off, n, j->msv, j->tlog);
            else if (j->codec == SHIM_HUF) r = HUF_compress2(c, j->slot, j->src + off, n, j->msv, j->tlog);
            else r = FSE_compressU16(c, j->slot, (const unsigned short*)(const void*)(j->src + off), n / 2, j->msv, j->tlog);
            j->csizes[b] = r;
        } else {
            size_t const cs = j->csizes[b];
            size_t r;
            if (cs == 0) { memcpy(j->out + off, j->src + off, n); r = n; }                 /* bench.c:393-397 */
            else if (cs == 1 && j->codec != SHIM_FSEU16) { memset(j->out + off, j->src[off], n); r = n; }  /* bench.c:398-402 */
            else if (j->codec == SHIM_FSE) r = FSE_decompress(j->out + off, n, c, cs);
            else if (j->codec == SHIM_HUF) r = HUF_decompress(j->out + off, n, c, cs);
            else {
                r = FSE_decompressU16((unsigned short*)(void*)(j->out + off), n / 2, c, cs);
                if (!FSE_isError(r)) r *= 2;
            }
            if (j->results) j->results[b] = r;
        }
    }
    return NULL;
}

static double run(job_t proto, size_t nBlocks, int nthreads)
{
    pthread_t* th;
    job_t* jobs;
    int t;
    double t0, t1;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > nBlocks && nBlocks) nthreads = (int)nBlocks;
    th = (pthread_t*)malloc(sizeof(*th) * (size_t)nthreads);
    jobs = (job_t*)malloc(sizeof(*jobs) * (size_t)nthreads);
    for (t = 0; t < nthreads; t++) {
        jobs[t] = proto;
        jobs[t].first = nBlocks * (size_t)t / (size_t)nthreads;
        jobs[t].last = nBlocks * (size_t)(t + 1) / (size_t)nthreads;
    }
    t0 = now_s();
    if (nthreads == 1) worker(&jobs[0]);
    else {
        for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
        for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    t1 = now_s();
    free(th); free(jobs);
    return t1 - t0;
}

/* number of blocks the way the shim splits: ceil(total / blockSize) (no trailing empty block) */
size_t refshim_nblocks(size_t total, size_t blockSize) { return (total + blockSize - 1) / blockSize; }

/* Compress every block of src[0..total) into cbuf + b*slot; csizes[b] = reference return value.
 * Returns elapsed wall-clock seconds. */
double refshim_compress_blocks(int codec, const void* src, size_t total, size_t blockSize,
                               void* cbuf, size_t slot, size_t* csizes,
                               unsigned maxSymbolValue, unsigned tableLog, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.codec = codec; j.decode = 0;
    j.src = (const unsigned char*)src; j.cbuf = (unsigned char*)cbuf;
    j.total = total; j.blockSize = blockSize; j.slot = slot; j.csizes = csizes;
    j.msv = maxSymbolValue; j.tlog = tableLog;
    return run(j, refshim_nblocks(total, blockSize), nthreads);
}

/* Decompress every block (raw / RLE handled as programs/bench.c does, from `orig`). */
double refshim_decompress_blocks(int codec, void* out, const void* orig, size_t total, size_t blockSize,
                                 const void* cbuf, size_t slot, const size_t* csizes,
                                 size_t* results, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.codec = codec; j.decode = 1;
    j.src = (const unsigned char*)orig; j.out = (unsigned char*)out;
    j.cbuf = (unsigned char*)(size_t)cbuf;
    j.total = total; j.blockSize = blockSize; j.slot = slot;
    j.csizes = (size_t*)(size_t)csizes; j.results = results;
    return run(j, refshim_nblocks(total, blockSize), nthreads);
}
