/* vecsim_b200.h — C-ABI of libvecsim_b200.so, the B200-native drop-in for the FLAT (brute-force)
 * slice of the reference's VecSim C API.
 *
 * Every VecSim* entry point below has the same name, argument order, argument meaning, ownership
 * rule and error behaviour as the reference declaration cited next to it (paths relative to
 * /root/reference/deps/VectorSimilarity/src/VecSim/), so RediSearch's callers —
 * src/vector_index.c, src/iterators/hybrid_reader.c, src/document.c:721, src/spec.c:3539 — link
 * against this library unchanged.  Enum values and struct layouts are ABI-identical to
 * vec_sim_common.h / query_results.h (checked by tests/test_abi_layout.py against golden
 * sizeof/offsetof tables taken from the reference headers).  The text of this header is written
 * from scratch; only names and layouts are shared, because they are the ABI.
 *
 * What differs from the reference is where the work happens: the corpus lives in HBM and every
 * distance/top-k computation is a hand-written sm_100a CUDA kernel.  There is no CPU fallback:
 * if no CUDA device is usable VecSimIndex_New returns NULL and logs through the log callback.
 *
 * VecSimB200_* entry points are extensions the reference does not have (batched queries, bulk
 * device ingest, shard merge); INTEGRATION.md shows where a RediSearch maintainer would call them.
 */
#ifndef VECSIM_B200_H
#define VECSIM_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Enums — numeric values follow vec_sim_common.h:60-117 and query_results.h:21-26.
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    VecSimType_FLOAT32 = 0,
    VecSimType_FLOAT64 = 1, /* not on the B200 path: VecSimIndex_New returns NULL */
    VecSimType_BFLOAT16 = 2,
    VecSimType_FLOAT16 = 3,
    VecSimType_INT8 = 4,
    VecSimType_UINT8 = 5,
    VecSimType_INT32 = 6,
    VecSimType_INT64 = 7
} VecSimType;

typedef enum {
    VecSimAlgo_BF = 0, /* the only algorithm this library implements */
    VecSimAlgo_HNSWLIB = 1,
    VecSimAlgo_TIERED = 2,
    VecSimAlgo_SVS = 3
} VecSimAlgo;

typedef enum { VecSimMetric_L2 = 0, VecSimMetric_IP = 1, VecSimMetric_Cosine = 2 } VecSimMetric;

typedef enum { VecSimOption_AUTO = 0, VecSimOption_ENABLE = 1, VecSimOption_DISABLE = 2 } VecSimOptionMode;
typedef enum { VecSimBool_TRUE = 1, VecSimBool_FALSE = 0, VecSimBool_UNSET = -1 } VecSimBool;

typedef enum { BY_SCORE = 0, BY_ID = 1, BY_SCORE_THEN_ID = 2 } VecSimQueryReply_Order;
typedef enum { VecSim_QueryReply_OK = 0, VecSim_QueryReply_TimedOut = 1 } VecSimQueryReply_Code;

#define VecSim_OK 0
typedef enum {
    VecSimParamResolver_OK = VecSim_OK,
    VecSimParamResolverErr_NullParam,
    VecSimParamResolverErr_AlreadySet,
    VecSimParamResolverErr_UnknownParam,
    VecSimParamResolverErr_BadValue,
    VecSimParamResolverErr_InvalidPolicy_NExits,
    VecSimParamResolverErr_InvalidPolicy_NHybrid,
    VecSimParamResolverErr_InvalidPolicy_NRange,
    VecSimParamResolverErr_InvalidPolicy_AdHoc_With_BatchSize,
    VecSimParamResolverErr_InvalidPolicy_AdHoc_With_EfRuntime
} VecSimResolveCode;

/* vec_sim_common.h:303-319 */
typedef enum {
    EMPTY_MODE = 0,
    STANDARD_KNN,
    HYBRID_ADHOC_BF,
    HYBRID_BATCHES,
    HYBRID_BATCHES_TO_ADHOC_BF,
    RANGE_QUERY
} VecSearchMode;

typedef enum { QUERY_TYPE_NONE = 0, QUERY_TYPE_KNN, QUERY_TYPE_HYBRID, QUERY_TYPE_RANGE } VecsimQueryType;

typedef enum { VecSim_WriteAsync = 0, VecSim_WriteInPlace = 1 } VecSimWriteMode;

typedef enum {
    VecSimSvsQuant_NONE = 0,
    VecSimSvsQuant_Scalar = 1,
    VecSimSvsQuant_4 = 4,
    VecSimSvsQuant_8 = 8,
    VecSimSvsQuant_4x4 = 4 | (4 << 8),
    VecSimSvsQuant_4x8 = 4 | (8 << 8),
    VecSimSvsQuant_4x8_LeanVec = 4 | (8 << 8) | (1 << 16),
    VecSimSvsQuant_8x8_LeanVec = 8 | (8 << 8) | (1 << 16)
} VecSimSvsQuantBits;

#define DEFAULT_BLOCK_SIZE 1024
#define VECSIM_POLICY_ADHOC_BF "adhoc_bf"
#define VECSIM_POLICY_BATCHES "batches"

typedef size_t labelType;      /* == RediSearch t_docId */
typedef unsigned int idType;   /* dense internal row id */

/* ------------------------------------------------------------------------------------------------
 * Parameter structs.  Only BFParams is interpreted; the other arms of the AlgoParams union are
 * declared so that sizeof(VecSimParams) and offsetof(VecSimParams, logCtx) match the reference
 * (vec_sim_common.h:142-268).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const char *name;
    size_t nameLen;
    const char *value;
    size_t valLen;
} VecSimRawParam;

typedef struct {
    VecSimType type;
    size_t dim;
    VecSimMetric metric;
    bool multi;             /* true: one label may own several vectors (JSON multi-value) */
    size_t initialCapacity; /* deprecated upstream; here: rows of HBM to reserve up front */
    size_t blockSize;       /* 0 -> DEFAULT_BLOCK_SIZE; growth granule of the label table */
} BFParams;

typedef struct {
    VecSimType type;
    size_t dim;
    VecSimMetric metric;
    bool multi;
    size_t initialCapacity;
    size_t blockSize;
    size_t M;
    size_t efConstruction;
    size_t efRuntime;
    double epsilon;
} HNSWParams;

typedef struct {
    VecSimType type;
    size_t dim;
    VecSimMetric metric;
    bool multi;
    size_t blockSize;
    VecSimSvsQuantBits quantBits;
    float alpha;
    size_t graph_max_degree;
    size_t construction_window_size;
    size_t max_candidate_pool_size;
    size_t prune_to;
    VecSimOptionMode use_search_history;
    size_t num_threads;
    size_t search_window_size;
    size_t search_buffer_capacity;
    size_t leanvec_dim;
    double epsilon;
} SVSParams;

typedef struct AsyncJob AsyncJob;
typedef void (*JobCallback)(AsyncJob *);
typedef int (*SubmitCB)(void *job_queue, void *index_ctx, AsyncJob **jobs, JobCallback *CBs, size_t jobs_len);
typedef struct VecSimParams VecSimParams;

typedef struct { size_t swapJobThreshold; } TieredHNSWParams;
typedef struct { char _placeholder; } TieredHNSWDiskParams;
typedef struct {
    size_t trainingTriggerThreshold;
    size_t updateTriggerThreshold;
    size_t updateJobWaitTime;
} TieredSVSParams;

typedef struct {
    void *jobQueue;
    void *jobQueueCtx;
    SubmitCB submitCb;
    size_t flatBufferLimit;
    VecSimParams *primaryIndexParams;
    union {
        TieredHNSWParams tieredHnswParams;
        TieredSVSParams tieredSVSParams;
        TieredHNSWDiskParams tieredHnswDiskParams;
    } specificParams;
} TieredIndexParams;

typedef union {
    HNSWParams hnswParams;
    BFParams bfParams;
    TieredIndexParams tieredParams;
    SVSParams svsParams;
} AlgoParams;

struct VecSimParams {
    VecSimAlgo algo; /* must be VecSimAlgo_BF */
    AlgoParams algoParams;
    void *logCtx; /* handed back as the first argument of the log callback */
};

/* Runtime (per query) parameters, vec_sim_common.h:283-338.  FLAT reads batchSize, searchMode and
 * timeoutCtx only. */
typedef struct { size_t efRuntime; double epsilon; } HNSWRuntimeParams;
typedef struct { size_t efRuntime; double epsilon; VecSimBool shouldRerank; } HNSWDiskRuntimeParams;
typedef struct {
    size_t windowSize;
    size_t bufferCapacity;
    VecSimOptionMode searchHistory;
    double epsilon;
} SVSRuntimeParams;

typedef struct {
    union {
        HNSWRuntimeParams hnswRuntimeParams;
        HNSWDiskRuntimeParams hnswDiskRuntimeParams;
        SVSRuntimeParams svsRuntimeParams;
    };
    size_t batchSize;
    VecSearchMode searchMode;
    void *timeoutCtx; /* passed to the timeout callback between kernel launches */
} VecSimQueryParams;

/* vec_sim_common.h:343-360 */
typedef struct {
    VecSimAlgo algo;
    VecSimMetric metric;
    VecSimType type;
    bool isMulti;
    bool isTiered;
    bool isDisk;
    size_t blockSize;
    size_t dim;
} VecSimIndexBasicInfo;

typedef struct {
    size_t memory; /* host bytes + HBM bytes owned by the index */
    size_t numberOfMarkedDeleted;
    size_t directHNSWInsertions;
    size_t flatBufferSize;
} VecSimIndexStatsInfo;

/* VecSimIndexDebugInfo (vec_sim_common.h:372-457): returned BY VALUE by VecSimIndex_DebugInfo, so the whole union has to keep
 * the reference's size and member offsets even though a FLAT index only fills commonInfo and bfInfo. */
typedef struct {
    VecSimIndexBasicInfo basicInfo;
    size_t indexSize;       /* current count of vectors */
    size_t indexLabelCount; /* current unique count of labels */
    uint64_t memory;
    VecSearchMode lastMode; /* the mode in which the last query ran */
} CommonInfo;
typedef struct { size_t M, efConstruction, efRuntime; double epsilon; size_t max_level, entrypoint, visitedNodesPoolSize,
                 numberOfMarkedDeletedNodes; } hnswInfoStruct;
typedef struct { char dummy; } bfInfoStruct;
typedef struct {
    VecSimSvsQuantBits quantBits;
    float alpha;
    size_t graphMaxDegree, constructionWindowSize, maxCandidatePoolSize, pruneTo;
    bool useSearchHistory;
    size_t numThreads, lastReservedThreads, numberOfMarkedDeletedNodes, searchWindowSize, searchBufferCapacity, leanvecDim;
    double epsilon;
} svsInfoStruct;
typedef struct HnswTieredInfo { size_t pendingSwapJobsThreshold; } HnswTieredInfo;
typedef struct SvsTieredInfo { size_t trainingTriggerThreshold, updateTriggerThreshold, updateJobWaitTime; bool indexUpdateScheduled; } SvsTieredInfo;
typedef struct {
    union { hnswInfoStruct hnswInfo; svsInfoStruct svsInfo; } backendInfo;
    union { HnswTieredInfo hnswTieredInfo; SvsTieredInfo svsTieredInfo; } specificTieredBackendInfo;
    CommonInfo backendCommonInfo;
    CommonInfo frontendCommonInfo;
    bfInfoStruct bfInfo;
    uint64_t management_layer_memory;
    VecSimBool backgroundIndexing;
    size_t bufferLimit;
} tieredInfoStruct;
typedef struct {
    CommonInfo commonInfo;
    union {
        bfInfoStruct bfInfo;
        hnswInfoStruct hnswInfo;
        svsInfoStruct svsInfo;
        tieredInfoStruct tieredInfo;
    };
} VecSimIndexDebugInfo;

/* Debug-info iterator (info_iterator.h:17-44).  Fields reported for FLAT: ALGORITHM, TYPE,
 * DIMENSION, METRIC, IS_MULTI_VALUE, IS_DISK, INDEX_SIZE, INDEX_LABEL_COUNT, MEMORY,
 * LAST_SEARCH_MODE, BLOCK_SIZE (brute_force.h:327-365). */
typedef struct VecSimDebugInfoIterator VecSimDebugInfoIterator;
typedef enum { INFOFIELD_STRING, INFOFIELD_INT64, INFOFIELD_UINT64, INFOFIELD_FLOAT64, INFOFIELD_ITERATOR } VecSim_InfoFieldType;
typedef union {
    double floatingPointValue;
    int64_t integerValue;
    uint64_t uintegerValue;
    const char *stringValue;
    VecSimDebugInfoIterator *iteratorValue;
} FieldValue;
typedef struct {
    const char *fieldName;
    VecSim_InfoFieldType fieldType;
    FieldValue fieldValue;
} VecSim_InfoField;

typedef void *(*allocFn)(size_t n);
typedef void *(*callocFn)(size_t nelem, size_t elemsz);
typedef void *(*reallocFn)(void *p, size_t n);
typedef void (*freeFn)(void *p);
typedef struct {
    allocFn allocFunction;
    callocFn callocFunction;
    reallocFn reallocFunction;
    freeFn freeFunction;
} VecSimMemoryFunctions;

typedef int (*timeoutCallbackFunction)(void *ctx); /* non-zero = expired */
typedef void (*logCallbackFunction)(void *ctx, const char *level, const char *message);

/* Opaque handles. */
typedef struct VecSimIndexInterface VecSimIndex;
typedef struct VecSimQueryResult VecSimQueryResult;
typedef struct VecSimQueryReply VecSimQueryReply;
typedef struct VecSimQueryReply_Iterator VecSimQueryReply_Iterator;
typedef struct VecSimBatchIterator VecSimBatchIterator;
typedef struct VecSimAdhocBfCtx VecSimAdhocBfCtx;

/* ------------------------------------------------------------------------------------------------
 * Index lifetime and mutation  (vec_sim.h:28-71, vec_sim.cpp:213-236)
 * ---------------------------------------------------------------------------------------------- */
/* vec_sim.h:28.  NULL if params->algo != BF, the type is unsupported, or no CUDA device. */
VecSimIndex *VecSimIndex_New(const VecSimParams *params);
/* vec_sim.h:36,49: host-side estimate, same formula family as brute_force_factory.cpp:96-135. */
size_t VecSimIndex_EstimateInitialSize(const VecSimParams *params);
size_t VecSimIndex_EstimateElementSize(const VecSimParams *params);
/* vec_sim.h:55 */
void VecSimIndex_Free(VecSimIndex *index);
/* vec_sim.h:65.  Blob is copied (cosine: normalised copy, preprocessors.h:49-146).  Single-value
 * index: an existing label is overwritten in place and 0 is returned, else 1. */
int VecSimIndex_AddVector(VecSimIndex *index, const void *blob, size_t label);
/* vec_sim.h:73.  Swap-delete: the last row moves into the hole (brute_force.h:196-224). */
int VecSimIndex_DeleteVector(VecSimIndex *index, size_t label);
/* vec_sim.h:116 */
size_t VecSimIndex_IndexSize(VecSimIndex *index);

/* ------------------------------------------------------------------------------------------------
 * Queries
 * ---------------------------------------------------------------------------------------------- */
/* vec_sim.h:143 / brute_force.h:243-291.  k nearest by (score asc, label asc); k==0 -> empty;
 * k>size -> size results.  order BY_ID sorts the reply by label (vec_sim.cpp:353-355). */
VecSimQueryReply *VecSimIndex_TopKQuery(VecSimIndex *index, const void *queryBlob, size_t k,
                                        VecSimQueryParams *queryParams, VecSimQueryReply_Order order);
/* vec_sim.h:159 / brute_force.h:293-326.  All rows with score <= (float)radius.  Returns NULL and
 * logs for radius < 0 or an order other than BY_ID/BY_SCORE (the reference throws,
 * vec_sim.cpp:362-367). */
VecSimQueryReply *VecSimIndex_RangeQuery(VecSimIndex *index, const void *queryBlob, double radius,
                                         VecSimQueryParams *queryParams, VecSimQueryReply_Order order);
/* vec_sim.h:91 / brute_force_single.h:200-212.  blob must already be normalised for cosine.
 * NaN if the label is absent; multi-value: min over the label's vectors. */
double VecSimIndex_GetDistanceFrom_Unsafe(VecSimIndex *index, size_t label, const void *blob);
/* vec_sim.h:229 / brute_force.h:380-451 (decision tree reproduced exactly; sets lastMode). */
bool VecSimIndex_PreferAdHocSearch(VecSimIndex *index, size_t subsetSize, size_t k, bool initial_check);
/* vec_sim.h:128 / vec_sim.cpp:270-343 */
VecSimResolveCode VecSimIndex_ResolveParams(VecSimIndex *index, VecSimRawParam *rparams, int paramNum,
                                            VecSimQueryParams *qparams, VecsimQueryType query_type);

/* Batch iterator, vec_sim.h:205 + query_results.h:115-138 / bf_batch_iterator.h.  The first Next
 * runs one full scan into an HBM score array; every Next returns the next-best n_results. */
VecSimBatchIterator *VecSimBatchIterator_New(VecSimIndex *index, const void *queryBlob,
                                             VecSimQueryParams *queryParams);
VecSimQueryReply *VecSimBatchIterator_Next(VecSimBatchIterator *iterator, size_t n_results,
                                           VecSimQueryReply_Order order);
bool VecSimBatchIterator_HasNext(VecSimBatchIterator *iterator);
void VecSimBatchIterator_Reset(VecSimBatchIterator *iterator);
void VecSimBatchIterator_Free(VecSimBatchIterator *iterator);

/* Ad-hoc context, vec_sim.h:247-281.  The reference returns NULL for RAM indexes
 * (vec_sim_interface.h:203); here it is implemented because a per-label device round trip is the
 * wrong shape for a GPU: _New uploads the (normalised) query once, _GetExactDistances computes a
 * whole label batch with one gather kernel. */
VecSimAdhocBfCtx *VecSimIndex_AdhocBfCtx_New(VecSimIndex *index, const void *queryBlob);
void VecSimIndex_AdhocBfCtx_Free(VecSimAdhocBfCtx *ctx);
double VecSimIndex_AdhocBfCtx_GetDistanceFrom(VecSimAdhocBfCtx *ctx, size_t label);
void VecSimIndex_AdhocBfCtx_GetExactDistances(VecSimAdhocBfCtx *ctx, const size_t *labels,
                                              double *distances_out, size_t count);

/* Replies, query_results.h:31-101 / query_results.cpp:23-73. */
size_t VecSimQueryReply_Len(VecSimQueryReply *reply);
VecSimQueryReply_Code VecSimQueryReply_GetCode(VecSimQueryReply *reply);
void VecSimQueryReply_Free(VecSimQueryReply *reply);
VecSimQueryReply_Iterator *VecSimQueryReply_GetIterator(VecSimQueryReply *reply);
VecSimQueryResult *VecSimQueryReply_IteratorNext(VecSimQueryReply_Iterator *iterator);
bool VecSimQueryReply_IteratorHasNext(VecSimQueryReply_Iterator *iterator);
void VecSimQueryReply_IteratorReset(VecSimQueryReply_Iterator *iterator);
void VecSimQueryReply_IteratorFree(VecSimQueryReply_Iterator *iterator);
int64_t VecSimQueryResult_GetId(const VecSimQueryResult *item);   /* NULL -> INVALID_ID (UINT_MAX) */
double VecSimQueryResult_GetScore(const VecSimQueryResult *item); /* NULL -> NaN */

/* ------------------------------------------------------------------------------------------------
 * Blob helpers, info, process-wide hooks
 * ---------------------------------------------------------------------------------------------- */
/* vec_sim.h:100 / normalize_naive.h:23-88 (host arithmetic, bit-identical to the reference). */
void VecSim_Normalize(void *blob, size_t dim, VecSimType type);
/* vec_sim.h:113: dim*sizeof(type) (+4 for INT8/UINT8 cosine). */
size_t VecSimParams_GetQueryBlobSize(VecSimType type, size_t dim, VecSimMetric metric);

VecSimIndexBasicInfo VecSimIndex_BasicInfo(VecSimIndex *index);               /* vec_sim.h:178 */
VecSimIndexStatsInfo VecSimIndex_StatsInfo(VecSimIndex *index);               /* vec_sim.h:185 */
VecSimIndexDebugInfo VecSimIndex_DebugInfo(VecSimIndex *index);               /* vec_sim.h:170; FLAT: BruteForceIndex::debugInfo, brute_force.h:318-325 */
VecSimDebugInfoIterator *VecSimIndex_DebugInfoIterator(VecSimIndex *index);   /* vec_sim.h:193 */
size_t VecSimDebugInfoIterator_NumberOfFields(VecSimDebugInfoIterator *it);   /* info_iterator.h:50 */
bool VecSimDebugInfoIterator_HasNextField(VecSimDebugInfoIterator *it);
VecSim_InfoField *VecSimDebugInfoIterator_NextField(VecSimDebugInfoIterator *it);
void VecSimDebugInfoIterator_Free(VecSimDebugInfoIterator *it);

void VecSimTieredIndex_GC(VecSimIndex *index);                    /* no-op for FLAT, vec_sim.h:211 */
void VecSimTieredIndex_AcquireSharedLocks(VecSimIndex *index);    /* no-op, vec_sim.h:237 */
void VecSimTieredIndex_ReleaseSharedLocks(VecSimIndex *index);    /* no-op, vec_sim.h:239 */

void VecSim_SetMemoryFunctions(VecSimMemoryFunctions memoryfunctions);   /* vec_sim.h:288 */
void VecSim_SetTimeoutCallbackFunction(timeoutCallbackFunction callback); /* vec_sim.h:295 */
void VecSim_SetLogCallbackFunction(logCallbackFunction callback);        /* vec_sim.h:302 */
void VecSim_SetWriteMode(VecSimWriteMode mode);                           /* no-op, vec_sim.h:320 */
void VecSim_SetTestLogContext(const char *test_name, const char *test_type); /* vec_sim.h:303: names the log file of the reference's
                                                                                 test logger; kept and shown by the default log sink */
void VecSim_UpdateThreadPoolSize(size_t new_size);                        /* no-op, vec_sim.h:330 */
size_t VecSim_GetSharedMemory(void);                                      /* 0,    vec_sim.h:338 */

/* ------------------------------------------------------------------------------------------------
 * B200 extensions (no reference counterpart)
 * ---------------------------------------------------------------------------------------------- */
/* nq queries in one corpus pass.  queryBlobs: nq host blobs, qstride bytes apart.  Results are
 * written to out_labels / out_scores ([nq][k], row-major, ascending (score,label)); entries past
 * the number of hits carry label SIZE_MAX and score NaN.  Returns VecSim_QueryReply_OK /
 * _TimedOut, or -1 on a CUDA failure. */
int VecSimB200_TopKQueryBatch(VecSimIndex *index, const void *queryBlobs, size_t qstride, size_t nq,
                              size_t k, VecSimQueryParams *queryParams, size_t *out_labels,
                              double *out_scores);
/* Same, but queries and results are DEVICE pointers (fp/int data as the index type; labels are
 * int64, scores float).  Nothing crosses PCIe; the call only enqueues on `stream`
 * (a cudaStream_t cast to void*, NULL = the index's own stream) and returns. */
int VecSimB200_TopKQueryBatchDevice(VecSimIndex *index, const void *d_queries, size_t nq, size_t k,
                                    int64_t *d_out_labels, float *d_out_scores, void *stream);
/* Bulk ingest of n host blobs (stride bytes apart) with labels[i] (NULL -> label0+i).  Equivalent
 * to n VecSimIndex_AddVector calls on fresh labels, with one H2D transfer per staging buffer. */
int VecSimB200_AddVectors(VecSimIndex *index, const void *blobs, size_t stride, size_t n,
                          const size_t *labels, size_t label0);
/* Bulk ingest of n rows that are already in device memory in stored form (cosine: already
 * normalised; int8/uint8 cosine: norm appended), row pitch = stored row size. */
int VecSimB200_AddVectorsDevice(VecSimIndex *index, const void *d_rows, size_t n, size_t label0);
/* Make sure HBM for `rows` vectors is reserved (avoids regrowth copies). */
int VecSimB200_Reserve(VecSimIndex *index, size_t rows);
/* Block until all staged rows are resident in HBM. */
int VecSimB200_Flush(VecSimIndex *index);
/* Device pointer / pitch of the resident corpus (for tests and the bench's roofline maths). */
const void *VecSimB200_DeviceRows(VecSimIndex *index, size_t *row_pitch_bytes, size_t *rows);
/* Launch statistics since the last call with reset=true: number of this library's kernels
 * launched, and device microseconds of the dominant scan kernel measured with CUDA events on the
 * launching stream. */
typedef struct {
    uint64_t kernel_launches;
    uint64_t scan_launches;
    double scan_device_us;
    uint64_t scan_bytes; /* algorithmic bytes the timed scan launches covered */
} VecSimB200_Stats;
VecSimB200_Stats VecSimB200_GetStats(VecSimIndex *index, bool reset);
/* Copy stored-form rows [first_row, first_row + n_rows) (internal row order, after the storage preprocessor: what
 * DataBlocksContainer holds in the reference) from HBM into a tightly packed host buffer of n_rows * stored-size bytes.
 * For persistence (RDB save) and for checking the device's corpus against a CPU scan.  0 / -1. */
int VecSimB200_ReadRows(VecSimIndex *index, size_t first_row, size_t n_rows, void *host_dst);
/* G-way merge of per-shard top-k lists (the coordinator's knnPostProcess, src/module.c:3139-3176,
 * comparator VecSim utils/query_result_utils.h:19-23), on device: in = [G][nq][k] gathered
 * (score,label) pairs, out = [nq][k]. */
int VecSimB200_MergeShardTopK(const float *d_scores, const int64_t *d_labels, size_t G, size_t nq,
                              size_t k, float *d_out_scores, int64_t *d_out_labels, void *stream);
/* ---- sharded KNN across GPUs, the exchange inside the library -------------------------------------------
 * One process per GPU, every process holds one row shard in its own VecSimIndex.  A query batch is answered by: local scan
 * of the shard -> ONE ncclAllGather of the packed per-shard top-k (labels + scores of a rank in one block) over NVLink ->
 * G-way merge on device by (score, label) — the coordinator's knnPostProcess, src/module.c:3139-3176, comparator
 * VS/utils/query_result_utils.h:19-23.  NCCL is bound at run time (dlopen "libnccl.so.2", or VECSIM_B200_NCCL_LIB), so
 * single-GPU users never load it.  Bootstrap like any NCCL program: rank 0 calls _UniqueId, ships the 128 bytes to the
 * other ranks over whatever control channel the host has (RediSearch: the cluster bus), every rank calls _New. */
typedef struct VecSimB200_ShardGroup VecSimB200_ShardGroup;
/* The exchange block of one shard: [labels int64 x nq*k][scores float x nq*k] padded to 16 bytes; and the merge of G such
 * blocks laid out rank-major (what an all-gather leaves) — for hosts that move the blocks with their own transport. */
size_t VecSimB200_ShardBlockBytes(size_t nq, size_t k);
int VecSimB200_MergeShardBlocks(const void *d_blocks, size_t G, size_t nq, size_t k, float *d_out_scores, int64_t *d_out_labels,
                                void *stream);
int VecSimB200_ShardGroup_UniqueId(void *out128);                                        /* 0 = ok */
VecSimB200_ShardGroup *VecSimB200_ShardGroup_New(const void *id128, int rank, int world); /* collective; NULL on error */
void VecSimB200_ShardGroup_Free(VecSimB200_ShardGroup *g);
int VecSimB200_ShardGroup_Rank(const VecSimB200_ShardGroup *g);
int VecSimB200_ShardGroup_Size(const VecSimB200_ShardGroup *g);
/* Collective, enqueued on `stream`, nothing synchronised: d_queries = nq stored-form (normalised) query blobs on this
 * rank's device; every rank receives the merged [nq][k] labels (int64, -1 = empty) and distances. */
int VecSimB200_ShardGroup_TopKBatchDevice(VecSimB200_ShardGroup *g, VecSimIndex *shard, const void *d_queries, size_t nq, size_t k,
                                          int64_t *d_out_labels, float *d_out_scores, void *stream);
/* Collective, host buffers end to end (raw query blobs in as for VecSimIndex_TopKQuery; H2D, shard scan, all-gather,
 * merge, D2H inside the call).  Empty slots: label SIZE_MAX, score NaN.  Returns 0 / -1. */
int VecSimB200_ShardGroup_TopKBatch(VecSimB200_ShardGroup *g, VecSimIndex *shard, const void *queryBlobs, size_t qstride, size_t nq,
                                    size_t k, size_t *out_labels, double *out_scores);
/* The whole HybridIterator state machine of src/iterators/hybrid_reader.c in one call: mode choice (:668-691:
 * VecSimIndex_PreferAdHocSearch on the child's estimate unless qp->searchMode forces a policy), HYBRID_BATCHES with the
 * reference's batch-size formula (:400-404), the alternating merge of each BY_ID batch with the child (:140-169) and the
 * policy review that may restart in ad-hoc mode (:346-370, :430-438), and HYBRID_ADHOC_BF (:289-335) served by ONE fused
 * device call over the drained child docIds instead of a GetDistanceFrom round trip per document.  child_iterator: any
 * object with the reference's QueryIterator vtable (iterator_api.h:46-151) — a B200 AND / OR result or a host iterator; it
 * is NOT freed.  The k best are kept in the reference's heap order (cmpVecSimResByScore :35-44) and written ordered by
 * (score, docId).  *out_mode = VecSearchMode the query ended in (also VecSimIndex's LAST_SEARCH_MODE), *out_iterations =
 * batches run.  Returns a VecSimQueryReply_Code, or -1.  Exact score ties at the k-th place in ad-hoc mode resolve towards
 * the smaller docId (the reference's min-max heap evicts the smaller docId among tied worst entries). */
int VecSimB200_HybridTopK(VecSimIndex *index, const void *queryBlob, size_t k, void *child_iterator, VecSimQueryParams *queryParams,
                          size_t *out_labels, double *out_scores, size_t *out_count, int *out_mode, size_t *out_iterations);
/* Hybrid "filter AND KNN" in ad-hoc mode, fused: what HybridIterator does in HYBRID_ADHOC_BF mode
 * (src/iterators/hybrid_reader.c:289-335: read the child iterator's docIds in ascending order, GetDistanceFrom each,
 * keep the k best in a heap with strict `<` admission, skip NaN = deleted) — in one call.  doc_ids: the filter's
 * ascending docIds (= vector labels), on the host or on the device (ids_on_device != 0, e.g.
 * II_ResultSet_DeviceDocIds of the filter's AND/OR).  Writes up to k (label, distance) pairs ordered by
 * (distance asc, docId asc) and their number.  Returns 0; -2 if the index cannot serve it (multi-value index, or
 * labels too sparse for the dense docId -> row table) — the caller then stays on VecSimIndex_GetDistanceFrom_Unsafe. */
int VecSimB200_TopKFiltered(VecSimIndex *index, const void *queryBlob, size_t k, const uint32_t *doc_ids, size_t n, int ids_on_device,
                            size_t *out_labels, double *out_scores, size_t *out_count);
/* The same for nq hybrid queries in one call (k <= 128): queryBlobs[i] with the DEVICE-resident ascending docId list
 * d_doc_ids[i] of counts[i] entries (its filter's AND / OR result).  Every query's kernel chain is enqueued on its own stream
 * before anything is waited for.  out_labels / out_scores are [nq][k], out_counts[i] the entries written for query i. */
int VecSimB200_TopKFilteredBatch(VecSimIndex *index, const void *const *queryBlobs, size_t nq, size_t k, const uint32_t *const *d_doc_ids,
                                 const size_t *counts, size_t *out_labels, double *out_scores, size_t *out_counts);
/* Batched fp32 queries (cosine, and in mode 1 also L2 and raw inner product; nq >= 16, k <= 16, dim % 8 == 0,
 * >= 65536 rows) take a tcgen05 coarse
 * pass + exact rescoring from the fp32 rows + a per-query completeness proof, with the exact scan as
 * on-device fallback (csrc/coarse_tc.cu); results are identical either way.  mode: 0 = exact scans only,
 * 1 = coarse pass over an fp16 shadow copy of the rows (+50% HBM, built lazily by the first eligible
 * batch; once it is complete, single VecSimIndex_TopKQuery calls ride it too — 2.4 ms instead of 5.2 ms per query on
 * 10M x 768), 2 = TF32 coarse pass over the fp32 rows (no extra memory, ~4x slower than 1),
 * -1 = environment default (VECSIM_B200_COARSE, 1 unless set). */
void VecSimB200_SetCoarseMode(int mode);
/* Debug: after a VecSimB200_TopKQueryBatchDevice call, per-query flags (1 = answered by the tensor-core path on the first
 * tier's 24-entry candidate lists, 2 = by the second tier's 128-entry lists, 0 = fell back to the exact scan).  Returns -1
 * if the last batch did not take the coarse path. */
int VecSimB200_LastCoarseFlags(VecSimIndex *index, uint32_t *out_ok, size_t nq);
/* Debug: which route the last top-k query (single or batched) took: 0 = exact CUDA-core scan, 1 = tensor-core coarse pass + exact
 * rescoring + proof (fp32 cosine), 2 = tensor-core direct, k <= 128 (csrc/coarse_tc.cu): fp16 / bf16 corpora, inner
 * product or cosine — the fp32-accumulated products of the stored 16-bit values are the distances; int8 / uint8 corpora,
 * inner product or cosine — kind::i8 integer dot products are exact and the reference's float expression is applied
 * to them, bit-exact. */
int VecSimB200_LastBatchPath(VecSimIndex *index);
/* Library/ABI version and the SM arch the kernels were compiled for ("sm_100a"). */
const char *VecSimB200_Version(void);

#ifdef __cplusplus
}
#endif
#endif /* VECSIM_B200_H */
