/*
 * fse_oracle.h -- CPU oracle ("port") for the FSE / Huff0 32 KB block hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (finitestateentropy_b200/, include/)
 * may include, link or call this.  Allowed users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_ref.py) against
 * the reference library itself compiled from /root/reference (oracle/_ref/libfse_ref.so, see
 * oracle/Makefile), against the committed golden vectors in tests/golden/ (generated from that
 * reference build by tests/golden/make_golden.py) and against the known-answer values of
 * SURVEY.md section 6.3.  The reference has no golden compressed vectors of its own
 * (SURVEY.md section 4), so the compiled reference is the ground truth.
 *
 * All `file:line` citations are relative to /root/reference/.
 */
#ifndef FSE_ORACLE_H
#define FSE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error convention of lib/error_private.h:77-79 and lib/error_public.h:45-56 */
enum {
    ORC_OK = 0, ORC_GENERIC = 1, ORC_DST_TOO_SMALL = 2, ORC_SRC_WRONG = 3, ORC_CORRUPT = 4,
    ORC_TLOG_TOO_LARGE = 5, ORC_MSV_TOO_LARGE = 6, ORC_MSV_TOO_SMALL = 7, ORC_WKSP_TOO_SMALL = 8,
    ORC_MAXCODE = 9
};
#define ORC_ERROR(c) ((size_t)0 - (size_t)(c))
unsigned orc_is_error(size_t code);

/* ---- statistics and FSE tables ---- */
size_t   orc_hist_count(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize);
unsigned orc_optimal_tablelog(unsigned maxTableLog, size_t srcSize, unsigned maxSymbolValue, unsigned minus);
size_t   orc_fse_normalize(short* norm, unsigned tableLog, const unsigned* count, size_t total, unsigned maxSymbolValue);
size_t   orc_fse_ncount_bound(unsigned maxSymbolValue, unsigned tableLog);
size_t   orc_fse_write_ncount(void* dst, size_t dstCap, const short* norm, unsigned maxSymbolValue, unsigned tableLog);
size_t   orc_fse_read_ncount(short* norm, unsigned* maxSVPtr, unsigned* tableLogPtr, const void* src, size_t srcSize);
/* tables use the reference's binary layouts (lib/fse.h:295-296,483-486,565-575; lib/fseU16.c:78-82) */
size_t   orc_fse_build_ctable(uint32_t* ct, const short* norm, unsigned maxSymbolValue, unsigned tableLog);
size_t   orc_fse_build_dtable(uint32_t* dt, const short* norm, unsigned maxSymbolValue, unsigned tableLog);
size_t   orc_fse_build_dtable_u16(uint32_t* dt, const short* norm, unsigned maxSymbolValue, unsigned tableLog);

/* ---- FSE stream codecs and block drivers ---- */
size_t orc_fse_encode(void* dst, size_t dstCap, const void* src, size_t srcSize, const uint32_t* ct);
size_t orc_fse_decode(void* dst, size_t dstCap, const void* cSrc, size_t cSrcSize, const uint32_t* dt);
size_t orc_fse_compress2(void* dst, size_t dstCap, const void* src, size_t srcSize, unsigned maxSymbolValue, unsigned tableLog);
size_t orc_fse_decompress(void* dst, size_t dstCap, const void* cSrc, size_t cSrcSize);
size_t orc_fse_compress_u16(void* dst, size_t dstCap, const uint16_t* src, size_t srcSize, unsigned maxSymbolValue, unsigned tableLog);
size_t orc_fse_decompress_u16(uint16_t* dst, size_t dstCap, const void* cSrc, size_t cSrcSize);

/* ---- Huff0 ---- */
/* CTable cell = {uint16 val; uint8 nbBits; uint8 pad}  (lib/huf_compress.c:106-109) */
size_t orc_huf_build_ctable(uint32_t* ctable, const unsigned* count, unsigned maxSymbolValue, unsigned maxNbBits);
size_t orc_huf_write_ctable(void* dst, size_t dstCap, const uint32_t* ctable, unsigned maxSymbolValue, unsigned huffLog);
size_t orc_huf_encode4x(void* dst, size_t dstCap, const void* src, size_t srcSize, const uint32_t* ctable);
size_t orc_huf_encode1x(void* dst, size_t dstCap, const void* src, size_t srcSize, const uint32_t* ctable);
size_t orc_huf_compress2(void* dst, size_t dstCap, const void* src, size_t srcSize, unsigned maxSymbolValue, unsigned huffLog);
size_t orc_huf_read_stats(uint8_t* weights, size_t hwSize, uint32_t* rankStats, uint32_t* nbSymbolsPtr,
                          uint32_t* tableLogPtr, const void* src, size_t srcSize);
/* DTable X1 = {maxTableLog,tableType,tableLog,reserved} + {byte,nbBits}[2^tableLog] (lib/huf_decompress.c:99-116) */
size_t orc_huf_read_dtable_x1(uint32_t* dtable, const void* src, size_t srcSize);
size_t orc_huf_read_dtable_x2(uint32_t* dtable, const void* src, size_t srcSize);   /* a18: double-symbol table image */
size_t orc_huf_decode4x1(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const uint32_t* dtable);
size_t orc_huf_decode1x1(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const uint32_t* dtable);
size_t orc_huf_decode4x2(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const uint32_t* dtable);   /* double-symbol table */
size_t orc_huf_decode1x2(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const uint32_t* dtable);
size_t orc_huf_decompress4x1(void* dst, size_t dstSize, const void* cSrc, size_t cSize);
size_t orc_huf_decompress4x2(void* dst, size_t dstSize, const void* cSrc, size_t cSize);
size_t orc_huf_decompress(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize);
unsigned orc_huf_select_decoder(size_t dstSize, size_t cSrcSize);

/* ---- measurement inputs and hashes ---- */
void     orc_probagen(void* buf, size_t size, double p);                              /* programs/probaGenerator.c:95-126 */
void     orc_gen_u16(uint16_t* buf, size_t nbSymbols, unsigned start, double p, uint32_t seed); /* programs/fuzzerU16.c:107-134 */
uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed);
uint32_t orc_xxh32(const void* data, size_t len, uint32_t seed);

/* block loops (the chunk loops of programs/bench.c:353-364 / :389-424) for timing the port; codec 0=FSE 1=HUF 2=FSEU16 */
size_t orc_compress_blocks(int codec, const void* src, size_t total, size_t blockSize, void* cbuf, size_t slot,
                           size_t* csizes, unsigned maxSymbolValue, unsigned tableLog);
size_t orc_decompress_blocks(int codec, void* out, const void* orig, size_t total, size_t blockSize,
                             const void* cbuf, size_t slot, const size_t* csizes, size_t* results);

#ifdef __cplusplus
}
#endif
#endif
