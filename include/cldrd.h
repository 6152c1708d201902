/* cldrd.h — C ABI of libcldrd.so: B200-native exhaustive inner-product top-k search.
 *
 * This is the drop-in boundary for CL-DRD's dense-retrieval search path.  The reference has no
 * FFI of its own: its seam is the subset of the `faiss` Python API that the retriever scripts
 * touch.  Every entry point below names the reference interface it replaces
 * (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding that a
 * maintainer of the reference would add.
 *
 * Conventions: plain C, no exceptions cross the boundary.  Every function returns 0 on success
 * or a negative CLDRD_E* code; cldrd_last_error() returns a thread-local message for the last
 * failure on the calling thread.  Pointers named *_host are host memory, *_dev are device
 * memory on the shard's device.  Handles are opaque and not thread-safe per handle (the
 * reference issues one search at a time: retriever/retrieval_utils.py:141-147).
 * There is NO CPU search path: without a CUDA device the compute entry points fail with
 * CLDRD_ECUDA.
 */
#ifndef CLDRD_H_
#define CLDRD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLDRD_ABI_VERSION 2

/* error codes */
#define CLDRD_OK        0
#define CLDRD_EINVAL   -1   /* bad argument (shape, dtype, k out of range, NULL) */
#define CLDRD_EIO      -2   /* file could not be opened / read / written */
#define CLDRD_EFORMAT  -3   /* not an IxMp/IxM2/IxFI index file, or metric != inner product */
#define CLDRD_ECUDA    -4   /* CUDA runtime / driver failure, or no device */
#define CLDRD_ENOMEM   -5   /* host or device allocation failed */
#define CLDRD_ESTATE   -6   /* call order violated (e.g. search before finalize) */

/* scan precision of a shard.  All modes return exact fp32 scores: the tensor-core scan is a
 * filter with a proven error band and every surviving candidate is re-scored in fp32 from the
 * fp32 rows.  They differ in what the scan streams from HBM. */
#define CLDRD_SCAN_SIMT_F32 0  /* fp32 FFMA tiles, no tensor cores; any d.  Correctness anchor.   */
#define CLDRD_SCAN_TC_TF32  1  /* tcgen05 kind::tf32 straight off the fp32 rows (no second copy). */
#define CLDRD_SCAN_TC_F16   2  /* tcgen05 kind::f16 over an fp16 copy of the rows.                */
#define CLDRD_SCAN_TC_BF16  3  /* tcgen05 kind::f16 over a bf16 copy of the rows.                 */

#define CLDRD_MAX_K 2048       /* same limit as faiss' GPU flat index */
#define CLDRD_SEED_J 32        /* sample scores kept per query for the seeded threshold */

typedef struct cldrd_shard cldrd_shard;

const char* cldrd_last_error(void);
int         cldrd_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Index file  (replaces faiss.write_index / faiss.read_index:
 *   retriever/index_text.py:105, retriever/retrieve_top_passages.py:85,
 *   retriever/retrieve_top_queries.py:60).  Layout: SURVEY.md §8 a-2.  Host only, no CUDA.
 * ------------------------------------------------------------------------------------------- */

/* Parse the headers.  has_ids: 1 for IxMp / IxM2 wrappers, 0 for a bare IxFI.  idmap2: 1 for IxM2.
 * data_off / ids_off: byte offsets of the float32 payload and of the int64 id array. */
int cldrd_index_probe(const char* path, int64_t* ntotal, int32_t* d, int32_t* metric,
                      int32_t* has_ids, int32_t* idmap2, int64_t* data_off, int64_t* ids_off);

/* Write IxMp{IxFI} (ids != NULL; IxM2 when idmap2 != 0) or bare IxFI (ids == NULL). */
int cldrd_index_write(const char* path, const float* xb_host, const int64_t* ids_host,
                      int64_t n, int32_t d, int32_t idmap2);

/* Streaming writer used by the index builder (replaces the hold-everything-in-RAM flow of
 * retriever/index_text.py:86-105): begin reserves the layout for n rows, append writes rows at
 * their final offset, finish writes the id array and closes. */
typedef struct cldrd_index_writer cldrd_index_writer;
int cldrd_index_writer_begin(cldrd_index_writer** out, const char* path, int64_t n, int32_t d,
                             int32_t with_ids, int32_t idmap2);
/* Sharded build (one process per GPU encodes its slice of the collection, all write ONE file): the writer covers
 * rows [row0, row0 + nrows) of the n rows the file declares.  create != 0 (exactly one process, before the others
 * open the file) creates the file and writes the headers; that writer's finish also writes the id array of all n
 * rows.  The others pass create = 0 (the file must exist and declare the same n, d and layout: CLDRD_EFORMAT
 * otherwise) and NULL ids -- or the ids, when the file's creator is gone (a build that is being continued).  Rows land
 * at their final offsets, so the file is byte-identical to a single-process build whatever the order of the writes. */
int cldrd_index_writer_open_range(cldrd_index_writer** out, const char* path, int64_t n, int32_t d,
                                  int32_t with_ids, int32_t idmap2, int64_t row0, int64_t nrows,
                                  int32_t create);
/* Rows appended so far are on stable storage when this returns (fdatasync): the index builder calls it before it
 * records its progress, so that an interrupted build can be continued from the recorded row (the reference's build,
 * retriever/index_text.py:86-105, starts over: 2.5 h for 8.8 M passages, README.md:20). */
int cldrd_index_writer_sync(cldrd_index_writer* w);
int cldrd_index_writer_append(cldrd_index_writer* w, const float* rows_host, int64_t nrows);
int cldrd_index_writer_finish(cldrd_index_writer* w, const int64_t* ids_host /* n or NULL */);

/* pread a row range / id range of an index file into host memory. */
int cldrd_index_read_rows(const char* path, int64_t row0, int64_t nrows, float* out_host);
int cldrd_index_read_ids(const char* path, int64_t row0, int64_t nrows, int64_t* out_host);

/* ---------------------------------------------------------------------------------------------
 * Shard = a contiguous range of passage rows resident in one GPU's HBM
 * (replaces faiss.index_cpu_to_gpu / index_cpu_to_gpu_multiple(shard=True):
 *   retriever/retrieval_utils.py:155-184).
 * ------------------------------------------------------------------------------------------- */

/* row0: global row of the shard's first row (added to local rows in the results).
 * scan: one of CLDRD_SCAN_*.  Allocates nothing large until rows arrive. */
int cldrd_shard_create(cldrd_shard** out, int device, int64_t row0, int64_t nrows, int32_t d,
                       int32_t scan);
void cldrd_shard_destroy(cldrd_shard* s);

/* Three ways to populate the fp32 rows (exactly one must cover [0, nrows)):
 *  - upload: copy n rows from host memory into local rows [row_off, row_off+n)
 *  - load_file: pread this shard's rows [row0, row0+nrows) of an index file through a pinned
 *    double buffer (the payload starts at byte 82, never aligned, so it cannot be mapped in place)
 *  - adopt: borrow a device buffer of nrows*d floats that the caller keeps alive (zero copy). */
int cldrd_shard_upload(cldrd_shard* s, const float* rows_host, int64_t row_off, int64_t n);
int cldrd_shard_load_file(cldrd_shard* s, const char* path);
int cldrd_shard_adopt(cldrd_shard* s, const float* rows_dev);

/* External ids (faiss IndexIDMap::id_map; retriever/index_text.py:97).  ids_host: nrows int64
 * for this shard's rows, or NULL -> results carry global row numbers. */
int cldrd_shard_set_ids(cldrd_shard* s, const int64_t* ids_host);

/* Build the scan-side state: fp16/bf16 copy (if the scan mode needs one), row-norm bound,
 * TMA descriptors.  Must be called once after the rows are in place and before any search. */
int cldrd_shard_finalize(cldrd_shard* s, void* cuda_stream);

/* Properties. */
int64_t cldrd_shard_nrows(const cldrd_shard* s);
int32_t cldrd_shard_dim(const cldrd_shard* s);
int32_t cldrd_shard_scan(const cldrd_shard* s);
/* Bytes one full scan streams from HBM (the roofline's algorithmic bytes per index pass). */
int64_t cldrd_shard_scan_bytes(const cldrd_shard* s);

/* ---------------------------------------------------------------------------------------------
 * Search  (replaces index.search(x, k): retriever/retrieval_utils.py:135,143;
 *          duplicate call sites evaluation/utils.py:110,117).
 * ------------------------------------------------------------------------------------------- */

/* Device-resident search over one shard.
 *   q_dev         [nq, d] float32, C-contiguous, 16-byte aligned
 *   out_scores_dev[nq, k] float32, best first; -FLT_MAX padding when the shard has < k rows
 *   out_ids_dev   [nq, k] int64: external ids if translate_ids != 0 and ids were set, else
 *                 global rows (row0 + local row); -1 padding
 * Order: descending score, ties -> lower row.  Work is issued on `cuda_stream`; the call
 * synchronises that stream before returning (it reads back the overflow / watchdog flags), so
 * results are complete on return. */
int cldrd_search_dev(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k,
                     int32_t translate_ids, float* out_scores_dev, int64_t* out_ids_dev,
                     void* cuda_stream);

/* Sharded search; each rank owns one shard.  The NCCL-transport form, three steps per
 * query batch (the peer-memory form follows below):
 *   1. cldrd_sample_dev: dense scan of a small strided sample of this shard's rows; writes each
 *      query's best CLDRD_SEED_J sample scan scores to out_topj_dev [nq][CLDRD_SEED_J].
 *   2. all-gather those to [parts][nq][CLDRD_SEED_J] (NCCL, by the caller) and call
 *      cldrd_seed_from_samples: seed[q] = CLDRD_SEED_J-th best of the union, i.e. a scan-score
 *      threshold that sits near rank 3k of the WHOLE index.
 *   3. cldrd_search_dev_seeded on every shard with that seed: only rows above it are collected,
 *      re-scored and returned (top-k of the shard among them, -1 padded); eps2_out_dev receives
 *      2*eps per query.  After the caller has merged the shards' lists (cldrd_merge_w; cldrd.dist
 *      does it slice-wise on every rank after an all-to-all),
 *      cldrd_verify_seed flags every query whose k-th merged score does not clear seed + eps:
 *      those (rare) queries must be searched again with seed_dev == NULL.
 * Any seed is safe: it only decides how much work the filter does, never the result. */
int cldrd_sample_dev(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k, float* out_topj_dev,
                     void* cuda_stream);
int cldrd_seed_from_samples(int device, const float* topj_dev, int32_t parts, int64_t nq,
                            float* seed_out_dev, void* cuda_stream);
int cldrd_search_dev_seeded(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k,
                            int32_t translate_ids, const float* seed_dev, float* out_scores_dev,
                            int64_t* out_ids_dev, float* eps2_out_dev, void* cuda_stream);
int cldrd_verify_seed(int device, const float* scores_dev, int64_t nq, int32_t k,
                      const float* seed_dev, const float* eps2_dev, int32_t* fail_dev,
                      void* cuda_stream);
/* ---------------------------------------------------------------------------------------------
 * Sharded search on one node: one shard per GPU, exchanges over NVLink / NVSwitch peer memory
 * (what `index_cpu_to_gpu_multiple(..., shard=True)` + IndexShards' host-thread merge stand for in
 * retriever/retrieval_utils.py:174-182).  One process per GPU (blocks mapped with CUDA IPC) or one
 * process driving several GPUs (blocks addressed directly).
 *
 * Every rank owns an exchange block of cldrd_node_block_bytes(world, max_k, d) bytes with the same
 * layout; the peers' kernels store into it.  A batch of at most CLDRD_QUERY_BATCH replicated queries
 * is ONE asynchronous call per rank, cldrd_node_search_begin, which enqueues on the caller's stream:
 *   1. sample scan; the CLDRD_SEED_J best sample scores per query go to plane [rank] of every rank's
 *      sample buffer (peer stores by the kernel); barrier; every rank derives the same J levels per query
 *      from the union, best first; the last level seeds the filter threshold (a scan score near rank 3.5 k
 *      of the WHOLE index);
 *   2. fused scan + filter + select over the shard; how many of its candidates scan at or above each
 *      level goes to plane [rank] of every rank's count buffer; barrier;
 *   3. re-score: T = the highest level that >= k rows of the whole index reach (summed counts);
 *      candidates scanning below T - 2*eps are dropped unscored (k rows scanning >= T put the exact
 *      k-th score above T - eps, so a top-k row scans above T - 2*eps); the rest are scored in fp32,
 *      sorted and stored as u64 keys (score, GLOBAL row; the valid prefix of the list plus its length)
 *      straight into the key planes of the rank that merges the query (query i of the batch belongs to
 *      rank i / ceil(nq / world)); barrier;
 *   4. every rank merges its slice (a key's merged rank = its place in its own sorted list + one binary
 *      search per other list; same key order as the single-shard search: bit-identical result),
 *      checks the seed (k-th merged score >= seed + eps), applies id_map and stores the rows through
 *      out_scores / out_ids: any memory this GPU can address -- its own, the collecting rank's result
 *      buffer (cldrd_node_result_ptrs) or page-locked host memory; barrier;
 *   5. a status record (queries that have to be searched again: seed missed, survivor overflow) goes to
 *      page-locked memory; cldrd_node_search_end waits for the batch and returns it.
 * The barriers are flag kernels over the same peer memory; a rank that never arrives ends in CLDRD_ECUDA
 * after CLDRD_BARRIER_TIMEOUT_MS (default 20 000), never in a hung GPU.  No host synchronisation and no
 * NCCL call inside a seeded batch; up to 4 batches may be in flight per node (begin ... begin, end ... end),
 * every rank issuing the same sequence of calls.  seeded = 0 (small shards, retry of raised queries) runs
 * the progressive scheme instead of steps 1-2 and re-scores every candidate.
 * out_rows_dev (optional, int32 [nq]): output row of batch query i (default i).  q_dev, out_* and
 * out_rows_dev must stay valid until the batch has ended. */
#define CLDRD_MAX_PEERS 16
#define CLDRD_QUERY_BATCH 8192
typedef struct cldrd_node cldrd_node;
/* d > 0 adds a query buffer [CLDRD_QUERY_BATCH][d] float32 to the block (cldrd_node_spread_queries); 0 = none */
int64_t cldrd_node_block_bytes(int32_t world, int32_t max_k, int32_t d);
int  cldrd_node_create(cldrd_node** out, int device, int32_t world, int32_t rank, int32_t max_k, int32_t d);
/* CLDRD_PEER_HANDLE_BYTES bytes to hand to the other processes (any byte transport) */
int  cldrd_node_handle(const cldrd_node* n, void* out_handle);
void* cldrd_node_block(const cldrd_node* n);
/* make rank peer_rank's block addressable: `handle` from its process (CUDA IPC), or, inside one process, its
 * cldrd_node_block pointer and device (peer access is enabled when the devices differ) */
int  cldrd_node_attach(cldrd_node* n, int32_t peer_rank, const void* handle, void* ptr, int32_t peer_device);
/* unmap the peers' blocks (every rank detaches before anybody destroys) */
int  cldrd_node_detach(cldrd_node* n);
void cldrd_node_destroy(cldrd_node* n);
/* result buffers [CLDRD_QUERY_BATCH][max_k] float32 / int64 inside rank owner_rank's block, as this process
 * addresses them: what the ranks pass as out_scores / out_ids to collect a batch on one GPU */
int  cldrd_node_result_ptrs(const cldrd_node* n, int32_t owner_rank, void** scores, void** ids);
int  cldrd_node_search_begin(cldrd_shard* s, cldrd_node* n, const float* q_dev, int64_t nq, int32_t k,
                             int32_t seeded, float* out_scores, int64_t* out_ids,
                             const int32_t* out_rows_dev, const int64_t* id_map_dev, void* cuda_stream);
/* Replicated queries from HOST memory without `world` uploads of the same bytes (index.search(x, k) hands over host
 * arrays: retriever/retrieval_utils.py:135): every rank copies only rows [row0, row0 + nrows) of the batch -- its 1/world
 * -- host -> cldrd_node_query_ptr() + row0*d floats, then calls cldrd_node_spread_queries, which stores that part into
 * every other rank's query buffer over NVLink and ends in a barrier; afterwards cldrd_node_query_ptr() holds the whole
 * batch on every rank and is what the rank passes as q_dev.  All ranks call it, before the batch's search_begin. */
int  cldrd_node_query_ptr(const cldrd_node* n, void** q_dev);
int  cldrd_node_spread_queries(cldrd_shard* s, cldrd_node* n, int64_t row0, int64_t nrows, void* cuda_stream);
/* Output sets: several result buffers registered once per rank (count <= CLDRD_MAX_OUT_SETS; scores[j] / ids[j] are set
 * j's float32 / int64 [rows][k] arrays as THIS rank's GPU addresses them, e.g. sub-blocks of one shared page-locked
 * host mapping), so that the rank that hands results to the caller can pick, batch by batch, a buffer the caller no
 * longer references -- `index.search` returns arrays the caller owns -- without a host round trip to the
 * other ranks: cldrd_node_search_begin_set is cldrd_node_search_begin with the output given as (set, first row);
 * exactly one rank passes out_select >= 0, which a kernel publishes in every rank's block ahead of the merge; the
 * others pass -1 and follow. */
#define CLDRD_MAX_OUT_SETS 8
int  cldrd_node_set_outputs(cldrd_node* n, int32_t count, void* const* scores, void* const* ids);
int  cldrd_node_search_begin_set(cldrd_shard* s, cldrd_node* n, const float* q_dev, int64_t nq, int32_t k,
                                 int32_t seeded, int32_t out_select, int64_t out_row0,
                                 const int32_t* out_rows_dev, const int64_t* id_map_dev, void* cuda_stream);
/* Oldest batch in flight: waits for it, fills cldrd_shard_last_stats / last_scan_time, returns how many of
 * its queries have to be searched again (the same on every rank) and their batch indices, ascending. */
int  cldrd_node_search_end(cldrd_shard* s, cldrd_node* n, int32_t* nfail_out, int32_t* fail_idx_out,
                           int32_t cap);
/* How cldrd_node_search_end waits: 0 (default) spins on the batch's event (lowest latency: one batch, nothing else to
 * do); 1 polls it every 100 us, leaving the core to host work that runs beside a long search (run-file writer). */
int  cldrd_node_set_wait_mode(cldrd_node* n, int32_t mode);
/* device milliseconds of the last ended batch: [0] prep + sample + barrier + levels, [1] scan + select,
 * [2] counts + barrier + re-score/scatter, [3] barrier + merge/store, [4] barrier + status,
 * [5] device idle on this stream between the end of the previous batch and the start of this one,
 * [6] the part of [2] before the re-score starts (count kernel + waiting for the slowest rank's scan) */
int  cldrd_node_phase_ms(const cldrd_node* n, double out[7]);

/* Device memory that other processes of the node can map (the node blocks above are made of it).
 * cldrd_peer_alloc: cudaMalloc + an opaque CLDRD_PEER_HANDLE_BYTES handle to hand to the peers (any byte transport;
 * cldrd.dist sends them through its process group); cldrd_peer_open maps a peer's handle into this process
 * (`device` = the opening process' GPU; needs peer access between the two GPUs);
 * cldrd_peer_copy is a stream-ordered copy between any two pointers CUDA knows (own or mapped device
 * memory, page-locked host memory). */
#define CLDRD_PEER_HANDLE_BYTES 72
int cldrd_peer_alloc(int device, int64_t nbytes, void** out_ptr, void* out_handle);
int cldrd_peer_free(int device, void* ptr);
int cldrd_peer_open(int device, const void* handle, void** out_ptr);
int cldrd_peer_close(int device, void* ptr);
int cldrd_peer_copy(int device, void* dst, const void* src, int64_t nbytes, void* cuda_stream);

/* The error band uses the largest row norm of the index: shards of one index must agree on it
 * (all-reduce MAX of cldrd_shard_norm_bound, then cldrd_shard_set_norm_bound on every shard). */
int cldrd_shard_norm_bound(const cldrd_shard* s, float* out);
int cldrd_shard_set_norm_bound(cldrd_shard* s, float bound);

/* Host-buffer search: the call the faiss-shaped wrapper makes.  Copies q in through pinned
 * staging, searches, copies D and I out; synchronous on return like faiss. */
int cldrd_search_host(cldrd_shard* s, const float* q_host, int64_t nq, int32_t k,
                      float* out_scores_host, int64_t* out_ids_host);

/* Page-locked host memory for result buffers: cldrd_search_host writes D and I straight into
 * buffers obtained here (the DMA engine's only copy); any other host buffer is served through
 * an internal pinned staging buffer plus one memcpy. */
int  cldrd_host_alloc(void** out, int64_t nbytes);
void cldrd_host_free(void* p);
/* Page-lock memory the caller already owns (e.g. a shared-memory mapping that several ranks of one
 * node write their slice of the results into, each over its own PCIe link). */
int  cldrd_host_register(void* p, int64_t nbytes);
int  cldrd_host_unregister(void* p);
/* The address under which kernels on `device` store into page-locked host memory obtained from
 * cldrd_host_alloc / cldrd_host_register (the merge kernels of a sharded search write their slice of the
 * result straight into the caller's host arrays over their own PCIe link). */
int  cldrd_host_device_ptr(int device, void* host_ptr, void** out_dev_ptr);

/* Merge per-shard candidate lists (replaces faiss IndexShards' CPU merge_knn_results behind
 * retriever/retrieval_utils.py:176-182).  Inputs are [parts][nq][k] device arrays of scores and
 * GLOBAL rows as produced by cldrd_search_dev(translate_ids=0) on each shard and gathered to
 * one device (NCCL all_gather / gather by the caller).  id_map_dev: optional int64[ntotal]
 * global-row -> external-id table applied to the merged rows (NULL = keep rows). */
int cldrd_merge(int device, const float* scores_dev, const int64_t* rows_dev, int32_t parts,
                int64_t nq, int32_t k, const int64_t* id_map_dev, float* out_scores_dev,
                int64_t* out_ids_dev, void* cuda_stream);
/* Same with input lists of width w != k ([parts][nq][w]): a seeded sharded search returns far
 * fewer than k valid rows per shard, so callers gather only the first w columns. */
int cldrd_merge_w(int device, const float* scores_dev, const int64_t* rows_dev, int32_t parts,
                  int64_t nq, int32_t w, int32_t k, const int64_t* id_map_dev,
                  float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream);

/* Same with planes that hold more rows than are merged: input [parts][plane_rows][w], the first
 * nq <= plane_rows rows of every plane are merged (lists received slice-wise from an all-to-all:
 * plane_rows = slice, nq = the queries this rank owns). */
int cldrd_merge_planes(int device, const float* scores_dev, const int64_t* rows_dev, int32_t parts,
                       int64_t plane_rows, int64_t nq, int32_t w, int32_t k,
                       const int64_t* id_map_dev, float* out_scores_dev, int64_t* out_ids_dev,
                       void* cuda_stream);

/* Per-search statistics of the last cldrd_search_* call on this shard (for tests / bench):
 * stats[0] kernel launches, [1] index chunks scanned, [2] queries sent to the dense fallback,
 * [3] total candidates re-scored, [4] total survivors pushed by the fused filter,
 * [5] max candidate-list length, [6] tcgen05 tiles executed, [7] in-kernel exact compactions. */
int cldrd_shard_last_stats(const cldrd_shard* s, int64_t stats[8]);

/* Scan-kernel timing for the roofline line of bench.py: when on, every scan launch is bracketed
 * by CUDA events on the launching stream; after a search, last_scan_time returns the summed
 * device time of the scan kernels of that search and how many launches it covers. */
int cldrd_shard_set_profiling(cldrd_shard* s, int32_t on);
int cldrd_shard_last_scan_time(const cldrd_shard* s, double* scan_ms, int64_t* scan_launches);
/* Per-launch detail of the same: fills ms[i] / rows[i] (index rows scanned by launch i) for up to
 * `cap` launches and returns how many were written. */
int cldrd_shard_last_scan_launches(const cldrd_shard* s, double* ms, int64_t* rows, int32_t cap);

/* Where the tensor pipe idles: cycles the MMA-issuing thread of the filter scans spent waiting,
 * summed over CTAs since the last reset (profiling on): out[0] operands (TMA/L2), out[1] the
 * epilogue handing back a TMEM stage, out[2] the next work-unit id, out[3] issuer lifetime. */
int cldrd_shard_wait_cycles(cldrd_shard* s, uint64_t out[4], int32_t reset);

/* Debug / test hook: run only the scan kernel in dense mode and return the raw scan scores
 * (approximate for the tensor-core modes) of queries [0,nq) against local rows
 * [row_begin, row_begin+nrows), nrows <= 8192.  out_dev: [nq, nrows] float32. */
int cldrd_scan_dense_dev(cldrd_shard* s, const float* q_dev, int64_t nq, int64_t row_begin,
                         int64_t nrows, float* out_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Run file  (replaces the regroup + writer loops: retriever/retrieve_top_passages.py:90-109,
 *            retriever/retrieve_top_queries.py:65-82).  Host only.
 * Writes "qid\tdocid\trank\tscore\n" per hit; score text is Python's repr(float(np.float32)).
 * Consecutive equal qids continue one rank sequence (the reference's dict regroup); hits with
 * id -1 are written as the reference would write them.  append != 0 opens with "a".
 * lines_written (optional) receives the number of lines. */
int cldrd_write_run(const char* path, const int64_t* qids, const float* scores,
                    const int64_t* ids, int64_t nq, int32_t k, int32_t append,
                    int64_t* lines_written);

/* The same with an explicit number of formatting threads (0 = CLDRD_WRITER_THREADS, else all host
 * cores): rows are cut into ~4 MiB pieces, formatted in parallel and written with pwrite at
 * prefix-summed offsets, so the file is byte-identical to the single-threaded one.  At config 5
 * (502 939 x 200 = 100 M lines) the reference's Python loop (retrieve_top_passages.py:99-109) would
 * take minutes; one thread of this writer a few seconds. */
int cldrd_write_run_mt(const char* path, const int64_t* qids, const float* scores,
                       const int64_t* ids, int64_t nq, int32_t k, int32_t append, int32_t threads,
                       int64_t* lines_written);

/* Run-file reader (the other direction: what evaluation/retrieval_evaluator.py:46-63 and the curriculum
 * post-processing of the top-200 runs do line by line in Python; 100 M lines at config 5).  Every line is
 * `line.strip().split("\t")` as in the reference: 2 to 4 fields, fields 0 and 1 integers (qid, pid); rank and score are
 * not needed by either consumer (file order IS rank order) and are skipped.  Call with qids = pids = NULL to learn
 * *nlines, then with arrays of that capacity: line i's ids land in qids[i], pids[i].  threads: as cldrd_write_run_mt.
 * CLDRD_EFORMAT + *bad_line (0-based, optional) on the first line the reference's reader would reject. */
int cldrd_read_run(const char* path, int64_t* qids, int64_t* pids, int64_t capacity, int32_t threads,
                   int64_t* nlines, int64_t* bad_line);

/* Format one float exactly as the reference's f-string does; returns the length written
 * (buf must hold >= 32 bytes). */
int cldrd_format_score(float s, char* buf);

/* Self-check of the score formatter: the writer prints fp32 scores through a specialised exact routine (one
 * 64x64-bit product per score instead of a general double -> text conversion); this entry compares its text with
 * the general routine's on `count` fp32 bit patterns first, first + stride, ... (mod 2^32).  Returns the number of
 * patterns that differ (0 = identical; < 0 never), the first such pattern in *first_bad and how many patterns the
 * specialised routine handled itself in *fast_taken (both optional).  tools/check_score_text.py sweeps all 2^32. */
int64_t cldrd_format_score_selfcheck(uint32_t first, uint32_t stride, int64_t count, uint32_t* first_bad,
                                     int64_t* fast_taken);

#ifdef __cplusplus
}
#endif
#endif /* CLDRD_H_ */
