/*
 * fse_b200.h -- C-ABI of libfse_b200.so: the B200 (sm_100a) implementation of the FSE / Huff0 block
 * entropy-coding hot path of Cyan4973/FiniteStateEntropy.
 *
 * Plain C: pointers and sizes only.  Link with -lfse_b200 (the library carries its own CUDA runtime).
 * All `file:line` citations are relative to the reference tree (/root/reference).
 *
 * Conventions kept from the reference (lib/error_private.h:77-79, lib/error_public.h:45-56):
 *   - every function returns size_t; errors are (size_t)-code, test with FSE_isError()/HUF_isError();
 *   - compressors return 0 = "not compressible / does not fit, nothing stored" and 1 = "single symbol,
 *     use RLE" in-band (lib/fse.h:63-65, lib/huf.h:51-52; HUF also stores the byte in dst[0]);
 *   - tables are caller-owned `unsigned[]` with the reference's documented layouts and sizes
 *     (lib/fse.h:295-300, lib/huf.h:136-149).
 *
 * Tier 1 (FSEB200_*_batch) works on DEVICE memory and a CUDA stream, a whole batch per call.
 * Tier 2 (the reference's own names) works on HOST memory, one block per synchronous call, and is
 * implemented by running the same kernels with a batch of one -- a correct drop-in for unmodified
 * callers (programs/bench.c, the fuzzers), not the fast path.
 * There is no CPU fallback: with no usable CUDA device a data-path call aborts with a message.
 */
#ifndef FSE_B200_H
#define FSE_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Tier 1 -- batched entry points (extension).  They replace the per-chunk loops of the reference
 * harness, programs/bench.c:353-364 (compress) and :389-424 (decompress), by one launch.
 *
 * Geometry (programs/bench.c:530-548): a flat buffer of `total` uncompressed bytes is cut into
 * ceil(total/blockSize) blocks of `blockSize` bytes (the last one shorter); compressed block b lives in
 * the fixed slot dCBuf + b*slot and may use `slot` bytes (bench.c uses FSE_compressBound(blockSize));
 * dCSizes[b] is exactly what the reference's FSE_compress2 / HUF_compress2 / FSE_compressU16 returns
 * for that block (0, 1, size, or an error code).  Bytes of a slot beyond the returned size are scratch.
 * Decompression writes block b at dDst + b*blockSize and dResults[b] = regenerated size or error code;
 * blocks with dCSizes[b]==0 (stored raw) or ==1 (RLE) are regenerated from dOrig as bench.c:393-402
 * does when dOrig != NULL (HUF additionally handles cSize==origSize / cSize==1 itself, lib/huf.h:60-63).
 * For the U16 codec sizes are in BYTES (a block holds blockSize/2 symbols, bench.c:221).
 * All pointers are device pointers; `stream` is a cudaStream_t (NULL = default stream).
 * The compressed buffer must be readable for 32 bytes past the last block's compressed bytes (the decoders load aligned 16- / 32-byte
 * pieces; what lies beyond a block's own bytes is never interpreted) -- bench.c's own buffer of nbChunks * FSE_compressBound() has it.
 * Return value: 0, or an error code if the launch itself could not be made.
 * ------------------------------------------------------------------------------------------------ */
size_t FSEB200_HUF_compress_batch(void* dCBuf, size_t slot, size_t* dCSizes, const void* dSrc, size_t srcTotal,
                                  size_t blockSize, unsigned maxSymbolValue, unsigned tableLog, void* stream);
size_t FSEB200_HUF_decompress_batch(void* dDst, size_t dstTotal, size_t blockSize, const void* dCBuf, size_t slot,
                                    const size_t* dCSizes, size_t* dResults, const void* dOrig, void* stream);
size_t FSEB200_FSE_compress_batch(void* dCBuf, size_t slot, size_t* dCSizes, const void* dSrc, size_t srcTotal,
                                  size_t blockSize, unsigned maxSymbolValue, unsigned tableLog, void* stream);
size_t FSEB200_FSE_decompress_batch(void* dDst, size_t dstTotal, size_t blockSize, const void* dCBuf, size_t slot,
                                    const size_t* dCSizes, size_t* dResults, const void* dOrig, void* stream);
size_t FSEB200_FSEU16_compress_batch(void* dCBuf, size_t slot, size_t* dCSizes, const void* dSrc, size_t srcTotal,
                                     size_t blockSize, unsigned maxSymbolValue, unsigned tableLog, void* stream);
size_t FSEB200_FSEU16_decompress_batch(void* dDst, size_t dstTotal, size_t blockSize, const void* dCBuf, size_t slot,
                                       const size_t* dCSizes, size_t* dResults, const void* dOrig, void* stream);
size_t FSEB200_batch_blocks(size_t total, size_t blockSize);

/* Table reuse across blocks (lib/huf.h:191 per block; the shape programs/bench.c:610-633 and HUF_compress4X_repeat,
 * lib/huf_compress.c:664-712, reduce to when the previous table is kept): every block of the batch is coded with ONE
 * caller-supplied table -- dCTable = 256 HUF_CElt cells in DEVICE memory, the layout HUF_buildCTable produces.
 * dCSizes[b] = what HUF_compress4X_usingCTable returns for block b (6 + the four streams, no tree header; 0 if it cannot). */
size_t FSEB200_HUF_compress4X_usingCTable_batch(void* dCBuf, size_t slot, size_t* dCSizes, const void* dSrc, size_t srcTotal,
                                                size_t blockSize, const unsigned* dCTable, void* stream);

/* Tier 1b -- the same batches on HOST buffers (pinned or pageable): chunks are copied in, processed and
 * copied out on alternating CUDA streams so that PCIe transfers overlap the kernels.
 * codec: 0 = FSE, 1 = HUF, 2 = FSE-U16.  Synchronous.  Raw / RLE blocks (cSize 0 / 1) are regenerated from
 * hOrig on the host exactly as programs/bench.c:393-402 does. */
size_t FSEB200_compress_host(int codec, void* hCBuf, size_t slot, size_t* hCSizes, const void* hSrc, size_t srcTotal,
                             size_t blockSize, unsigned maxSymbolValue, unsigned tableLog);
size_t FSEB200_decompress_host(int codec, void* hDst, size_t dstTotal, size_t blockSize, const void* hCBuf, size_t slot,
                               const size_t* hCSizes, size_t* hResults, const void* hOrig);

/* Measurement inputs generated directly in device memory: byte i of the output equals byte
 * (streamOffset + i) of the reference generator's stream (programs/probaGenerator.c:95-126 with
 * probability p, seed 1; programs/fuzzerU16.c:107-134 with the given start / p / seed). */
size_t FSEB200_probagen(void* dDst, size_t nBytes, size_t streamOffset, double p, void* stream);
size_t FSEB200_genU16(void* dDst, size_t nSymbols, size_t streamOffset, unsigned start, double p, unsigned seed, void* stream);
int    FSEB200_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Tier 2 -- the reference API, host pointers.
 * ------------------------------------------------------------------------------------------------ */
/* lib/fse.h:43-105 */
unsigned    FSE_versionNumber(void);
size_t      FSE_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize);
size_t      FSE_compress2(void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                          unsigned maxSymbolValue, unsigned tableLog);
size_t      FSE_decompress(void* dst, size_t dstCapacity, const void* cSrc, size_t cSrcSize);
size_t      FSE_compressBound(size_t size);
unsigned    FSE_isError(size_t code);
const char* FSE_getErrorName(size_t code);
/* lib/fse.h:135-247 (detailed API) */
unsigned    FSE_optimalTableLog(unsigned maxTableLog, size_t srcSize, unsigned maxSymbolValue);
unsigned    FSE_optimalTableLog_internal(unsigned maxTableLog, size_t srcSize, unsigned maxSymbolValue, unsigned minus);
size_t      FSE_normalizeCount(short* normalizedCounter, unsigned tableLog, const unsigned* count,
                               size_t srcSize, unsigned maxSymbolValue);
size_t      FSE_NCountWriteBound(unsigned maxSymbolValue, unsigned tableLog);
size_t      FSE_writeNCount(void* buffer, size_t bufferSize, const short* normalizedCounter,
                            unsigned maxSymbolValue, unsigned tableLog);
size_t      FSE_readNCount(short* normalizedCounter, unsigned* maxSymbolValuePtr, unsigned* tableLogPtr,
                           const void* rBuffer, size_t rBuffSize);
unsigned*   FSE_createCTable(unsigned maxSymbolValue, unsigned tableLog);
void        FSE_freeCTable(unsigned* ct);
size_t      FSE_buildCTable(unsigned* ct, const short* normalizedCounter, unsigned maxSymbolValue, unsigned tableLog);
unsigned*   FSE_createDTable(unsigned tableLog);
void        FSE_freeDTable(unsigned* dt);
size_t      FSE_buildDTable(unsigned* dt, const short* normalizedCounter, unsigned maxSymbolValue, unsigned tableLog);
/* lib/hist.h:30-75 */
size_t      HIST_count(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize);
unsigned    HIST_isError(size_t code);
size_t      HIST_count_wksp(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize,
                            void* workSpace, size_t workSpaceSize);
size_t      HIST_countFast(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize);
size_t      HIST_countFast_wksp(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize,
                                void* workSpace, size_t workSpaceSize);
unsigned    HIST_count_simple(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize);
/* lib/huf.h:54-98 */
size_t      HUF_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize);
size_t      HUF_compress2(void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                          unsigned maxSymbolValue, unsigned tableLog);
size_t      HUF_compress4X_wksp(void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                                unsigned maxSymbolValue, unsigned tableLog, void* workSpace, size_t wkspSize);
size_t      HUF_decompress(void* dst, size_t originalSize, const void* cSrc, size_t cSrcSize);
size_t      HUF_compressBound(size_t size);
unsigned    HUF_isError(size_t code);
const char* HUF_getErrorName(size_t code);
/* lib/huf.h:155-280 (static-linking tier used by fullbench / the north star).
 * HUF_CElt is {U16 val; BYTE nbBits} in a 4-byte cell (lib/huf_compress.c:106-109); HUF_DTable is U32[]
 * (lib/huf.h:144-149). */
unsigned    HUF_optimalTableLog(unsigned maxTableLog, size_t srcSize, unsigned maxSymbolValue);
size_t      HUF_buildCTable(unsigned* CTable, const unsigned* count, unsigned maxSymbolValue, unsigned maxNbBits);
size_t      HUF_writeCTable(void* dst, size_t maxDstSize, const unsigned* CTable, unsigned maxSymbolValue, unsigned huffLog);
size_t      HUF_readStats(unsigned char* huffWeight, size_t hwSize, unsigned* rankStats, unsigned* nbSymbolsPtr,
                          unsigned* tableLogPtr, const void* src, size_t srcSize);
size_t      HUF_readDTableX1(unsigned* DTable, const void* src, size_t srcSize);
size_t      HUF_readDTableX2(unsigned* DTable, const void* src, size_t srcSize);   /* double-symbol table image (huf_decompress.c:551-649) */
unsigned    HUF_selectDecoder(size_t dstSize, size_t cSrcSize);
size_t      HUF_decompress4X1(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize);
size_t      HUF_decompress4X2(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize);
/* Payload coding with a caller-built table image (lib/fse.h:222,247 ; lib/huf.h:191,203,277).  The images are the
 * reference's own ABI layouts (FSE_CTable fse.h:295,483-486 ; FSE_DTable fse.h:296,565-575 ; HUF_CElt huf_compress.c:106-109 ;
 * HUF_DTable single-symbol huf_decompress.c:101,116), e.g. as produced by FSE_buildCTable / FSE_buildDTable /
 * HUF_buildCTable / HUF_readDTableX1 above or by the CPU library.  Single synchronous calls on host buffers, <= 16 MiB. */
/* CTable inspection helpers (lib/huf.h:196-199,221) */
unsigned    HUF_getNbBits(const void* symbolTable, unsigned symbolValue);
size_t      HUF_estimateCompressedSize(const unsigned* CTable, const unsigned* count, unsigned maxSymbolValue);
int         HUF_validateCTable(const unsigned* CTable, const unsigned* count, unsigned maxSymbolValue);
/* constant-pattern tables for stored / single-symbol blocks (lib/fse.h:330-345) */
size_t      FSE_buildCTable_raw(unsigned* ct, unsigned nbBits);
size_t      FSE_buildCTable_rle(unsigned* ct, unsigned char symbolValue);
size_t      FSE_buildDTable_raw(unsigned* dt, unsigned nbBits);
size_t      FSE_buildDTable_rle(unsigned* dt, unsigned char symbolValue);
size_t      FSE_compress_usingCTable(void* dst, size_t dstCapacity, const void* src, size_t srcSize, const unsigned* ct);
size_t      FSE_decompress_usingDTable(void* dst, size_t dstCapacity, const void* cSrc, size_t cSrcSize, const unsigned* dt);
size_t      HUF_compress4X_usingCTable(void* dst, size_t dstSize, const void* src, size_t srcSize, const unsigned* CTable);
size_t      HUF_decompress4X_usingDTable(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable);
size_t      HUF_decompress4X1_usingDTable(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable);
size_t      HUF_decompress4X2_usingDTable(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable);
size_t      HUF_decompress1X2_usingDTable(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable);
/* single-stream Huff0 (lib/huf.h:288-320) */
size_t      HUF_compress1X(void* dst, size_t dstSize, const void* src, size_t srcSize, unsigned maxSymbolValue, unsigned tableLog);
size_t      HUF_compress1X_usingCTable(void* dst, size_t dstSize, const void* src, size_t srcSize, const unsigned* CTable);
size_t      HUF_decompress1X1(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize);
size_t      HUF_decompress1X_usingDTable(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable);
size_t      HUF_decompress1X1_usingDTable(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable);
/* lib/huf.h:194-208,291-300: table reuse, one block per call.  `repeat` points to a HUF_repeat (an int: none 0, check 1, valid 2) */
size_t      HUF_compress4X_repeat(void* dst, size_t dstSize, const void* src, size_t srcSize, unsigned maxSymbolValue, unsigned tableLog,
                                  void* workSpace, size_t wkspSize, unsigned* hufTable, int* repeat, int preferRepeat, int bmi2);
size_t      HUF_compress1X_repeat(void* dst, size_t dstSize, const void* src, size_t srcSize, unsigned maxSymbolValue, unsigned tableLog,
                                  void* workSpace, size_t wkspSize, unsigned* hufTable, int* repeat, int preferRepeat, int bmi2);
/* lib/huf.h:231,304,329,333 */
size_t      HUF_readCTable(unsigned* CTable, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize, unsigned* hasZeroWeights);
size_t      HUF_decompress1X2(void* dst, size_t dstSize, const void* cSrc, size_t cSrcSize);
size_t      HUF_decompress1X_usingDTable_bmi2(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable, int bmi2);
size_t      HUF_decompress4X_usingDTable_bmi2(void* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const unsigned* DTable, int bmi2);
/* lib/fseU16.h:75-79 */
size_t      FSE_compressU16(void* dst, size_t dstCapacity, const unsigned short* src, size_t srcSize,
                            unsigned maxSymbolValue, unsigned tableLog);
size_t      FSE_decompressU16(unsigned short* dst, size_t dstCapacity, const void* cSrc, size_t cSrcSize);
/* lib/fseU16.c:121-145 (not in fseU16.h; programs/fuzzerU16.c:257 declares it `extern`): histogram of 16-bit symbols on the GPU.
 * *maxSymbolValuePtr in: largest symbol `count` has room for; out: largest symbol present.  Returns the largest count. */
size_t      FSE_countU16(unsigned* count, unsigned* maxSymbolValuePtr, const unsigned short* src, size_t srcSize);

#ifdef __cplusplus
}
#endif
#endif /* FSE_B200_H */
