/*
 * mrag.h -- C ABI of the B200-native exact cosine top-k scan ("mrag").
 *
 * This is the drop-in boundary for ONE path of ananthlk/Mobius-RAG: the SQL shape
 *
 *     SELECT ..., 1 - (embedding_vec <=> :q) AS similarity
 *     FROM rag_published_embeddings WHERE <filters> AND embedding_vec IS NOT NULL
 *     ORDER BY embedding_vec <=> :q LIMIT :k
 *
 * issued by  app/services/vector_store.py:274-287  (PgVectorStore._search_async)
 * and by     app/services/corpus_search.py:1525-1536 (_vector_arm).
 * The reference has no FFI of its own for this path (it sends SQL to pgvector), so
 * each entry point below names the reference statement it stands in for.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (streams travel as void*).
 *   - every function returns 0 on success, a negative mrag_status otherwise;
 *     mrag_last_error() returns a thread-local message for the last failure.
 *   - the caller owns every input buffer; the library copies what it keeps before
 *     returning.  The index handle owns all device memory.
 *   - "row" is a position in the index (0-based, append order); the host shim keeps the
 *     row -> UUID / text hydration table (the reference hydrates in the same SELECT,
 *     corpus_search.py:621-640).
 *   - mrag_search* are re-entrant on one handle from several host threads (the reference
 *     runs up to 5 narrow searches concurrently, corpus_search_agent.py:794-797);
 *     append / tombstone take the writer side of the same lock.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with MRAG_ERR_CUDA.
 */
#ifndef MRAG_H_
#define MRAG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRAG_VERSION_MAJOR 0
#define MRAG_VERSION_MINOR 1

typedef enum mrag_status {
    MRAG_OK = 0,
    MRAG_ERR_ARG = -1,      /* bad argument (null, dim mismatch, k out of range ...) */
    MRAG_ERR_CUDA = -2,     /* CUDA runtime / driver failure, or no device */
    MRAG_ERR_OOM = -3,      /* device or host allocation failed / capacity exceeded */
    MRAG_ERR_STATE = -4     /* call not valid in this state */
} mrag_status;

/* storage type of the corpus in HBM */
typedef enum mrag_dtype {
    MRAG_F32 = 0,  /* float4 rows, as pgvector stores them (add_pgvector_columns.py:49) */
    MRAG_BF16 = 1  /* bf16 rows (round-to-nearest-even of the fp32 input); 1e-2 score mode */
} mrag_dtype;

/* largest k the fused scan+select handles in one pass; callers may ask for up to
 * MRAG_MAX_K (vector arm LIMIT = k*2*over_fetch <= 1600, corpus_search.py:1458,3297),
 * larger k go through the multi-round path inside mrag_search. */
#define MRAG_FUSED_K 128
#define MRAG_MAX_K 2048

/* code-space limits of the denormalised metadata columns */
#define MRAG_PAYER_WORDS 16   /* payer codes < 1024  */
#define MRAG_SMALL_WORDS 4    /* state / program / authority / source_type codes < 256 */
#define MRAG_TAG_WORDS 8      /* document d: and p: tag codes, one bit each, < 512 (lexicon has ~231 codes, corpus_search_lexicon.py:4) */
#define MRAG_CODE_NONE 0xFFFF /* "column is empty string" code for payer; 0xFF for the u8 columns */

/*
 * One row of rag_published_embeddings as the scan sees it (app/models.py:242-280):
 * the document-level columns are denormalised on every row, exactly like the table.
 * Strings are dictionary-coded by the host shim (mrag/vocab.py); equality on codes is
 * equality on strings.
 */
typedef struct mrag_rowmeta {
    uint32_t doc_idx;     /* dense index of document_id                        */
    uint16_t payer;       /* code of document_payer                            */
    uint8_t  state;       /* code of document_state                            */
    uint8_t  program;     /* code of document_program                          */
    uint8_t  authority;   /* code of document_authority_level                  */
    uint8_t  source_type; /* code of source_type                               */
    uint8_t  valid;       /* 0 <=> embedding_vec IS NULL (vector_store.py:263-266) */
    uint8_t  reserved;
} mrag_rowmeta;

/* which clauses of mrag_filter are active */
enum {
    MRAG_F_PAYER       = 1u << 0,  /* payer IN payer_any  OR (payer IN payer_alt_any AND state = alt_state)
                                      -- corpus_search.py:524-535 incl. the FL-Medicaid MCO union      */
    MRAG_F_STATE       = 1u << 1,  /* document_state = state_eq            corpus_search.py:536-538  */
    MRAG_F_PROGRAM     = 1u << 2,  /* document_program = program_eq        corpus_search.py:539-541  */
    MRAG_F_AUTHORITY   = 1u << 3,  /* document_authority_level = ..        corpus_search.py:542-544  */
    MRAG_F_SOURCE_TYPE = 1u << 4,  /* source_type = ..                     vector_store.py:156       */
    MRAG_F_DOC_EQ      = 1u << 5,  /* document_id = :document_id           vector_store.py:247-249   */
    MRAG_F_DOC_POOL    = 1u << 6,  /* document_id = ANY(:inc_ids)          corpus_search.py:546-558  */
    MRAG_F_TAG_STRICT  = 1u << 7,  /* OR(state IN, program IN, payer IN)   corpus_search.py:1478-1496 */
    MRAG_F_TAG_RELAXED = 1u << 8,  /* doc d/p tag bitset & tag_any != 0    corpus_search.py:1497-1510 */
    MRAG_F_DOC_POOL_HANDLE = 1u << 9 /* document_id = ANY(pool) with the pool resident on the device (mrag_pool_build):
                                      no UUID list is marshalled per search   corpus_search_agent.py:1762-1888 */
};

typedef struct mrag_filter {
    uint32_t flags;
    /* MRAG_F_PAYER */
    uint64_t payer_any[MRAG_PAYER_WORDS];
    uint64_t payer_alt_any[MRAG_PAYER_WORDS];
    uint16_t alt_state;                /* state code required with payer_alt_any */
    /* equality clauses (codes); a code that exists nowhere in the corpus matches no row */
    uint16_t state_eq;
    uint16_t program_eq;
    uint16_t authority_eq;
    uint16_t source_type_eq;
    uint16_t reserved0;
    uint32_t doc_eq;
    /* MRAG_F_DOC_POOL: host array of doc_idx values (copied by the call) */
    const uint32_t* doc_pool;
    int64_t n_doc_pool;
    /* MRAG_F_TAG_STRICT: one OR group over three code sets */
    uint64_t tag_state_any[MRAG_SMALL_WORDS];
    uint64_t tag_program_any[MRAG_SMALL_WORDS];
    uint64_t tag_payer_any[MRAG_PAYER_WORDS];
    /* MRAG_F_TAG_RELAXED: document tag bits, any-of */
    uint64_t tag_any[MRAG_TAG_WORDS];
    /* MRAG_F_DOC_POOL_HANDLE: a pool built by mrag_pool_build on the SAME index (excludes MRAG_F_DOC_POOL) */
    const struct mrag_pool* pool;
} mrag_filter;

typedef struct mrag_index mrag_index;

/* --- lifecycle ---------------------------------------------------------------------- */

/* CREATE TABLE .. embedding_vec vector(dim)   (add_pgvector_columns.py:56-65).
 * capacity = maximum number of rows this shard will ever hold (device memory is
 * reserved up front: capacity * round_up(dim,64) * sizeof(elem)). */
int mrag_create(mrag_index** out, int dim, int dtype, int device, int64_t capacity);
int mrag_destroy(mrag_index* idx);

/* --- write side --------------------------------------------------------------------- */

/* UPDATE .. SET embedding_vec = CAST('[..]' AS vector)  (embedding_worker.py:65-94,
 * publish.py:327-362).  rows = n*dim float32, HOST memory, row-major; meta = n entries.
 * Elements must be finite (pgvector rejects NaN/Inf on input). Returns the first new row
 * index through *first_row (may be NULL). */
int mrag_append(mrag_index* idx, const float* rows, int64_t n, const mrag_rowmeta* meta,
                int64_t* first_row);
/* same, rows already resident on the index's device (fp32, row-major, pitch = dim).
 * `stream` is a cudaStream_t or NULL. */
int mrag_append_device(mrag_index* idx, const void* d_rows_f32, int64_t n,
                       const mrag_rowmeta* meta, int64_t* first_row, void* stream);

/* document_tags.d_tags / p_tags key sets as bitsets (app/models.py:525-543):
 * bits = n_docs * MRAG_TAG_WORDS u64, doc-major; replaces docs [first_doc, first_doc+n_docs). */
int mrag_set_doc_tags(mrag_index* idx, int64_t first_doc, const uint64_t* bits, int64_t n_docs);

/* DELETE FROM .. WHERE document_id = :id  (publish.py:310-313, vector_store.py:101-104):
 * clears the valid bit of every row whose doc_idx matches. *n_rows (may be NULL) gets the count. */
int mrag_tombstone_doc(mrag_index* idx, uint32_t doc_idx, int64_t* n_rows);

/* Of the mrag_size() slots in use: rows that still exist (inserted, not deleted) and rows that still have a vector.
 * The shard is append-only -- a re-published document (publish.py:310-362: DELETE + INSERT) takes new slots -- so
 * size - live is what a rebuild (snapshot the live rows, load into a fresh index) reclaims. */
int mrag_live_rows(mrag_index* idx, int64_t* live_rows, int64_t* rows_with_vector);

int64_t mrag_size(const mrag_index* idx);      /* rows appended so far (incl. tombstoned) */
int64_t mrag_capacity(const mrag_index* idx);
int mrag_dim(const mrag_index* idx);
int mrag_index_dtype(const mrag_index* idx);
int mrag_device(const mrag_index* idx);

/* --- snapshot --------------------------------------------------------------------------
 * One shard <-> one file: rows, 1/|x|, metadata columns, NULL / live bitmaps, document tag sets, hybrid features.
 * The GPU copy is derived data (re-creatable from rag_published_embeddings); `user_version` carries
 * corpus_state.corpus_version (publish.py:314) so a stale snapshot can be recognised. mrag_load creates the index
 * on `device` with room for `capacity` rows (<= 0: exactly the rows in the file). */
int mrag_save(mrag_index* idx, const char* path, int64_t user_version);
int mrag_load(mrag_index** out, const char* path, int device, int64_t capacity, int64_t* user_version);

/* --- read side ---------------------------------------------------------------------- */

/* search options */
enum {
    MRAG_OPT_DEVICE_IO   = 1u << 0, /* q / scores / rows / counts are DEVICE pointers on the index's device */
    MRAG_OPT_FORCE_GEMV  = 1u << 1, /* pin the CUDA-core streaming kernel (testing / tuning) */
    MRAG_OPT_FORCE_MMA   = 1u << 2, /* pin the tcgen05 kernel (testing / tuning)             */
    MRAG_OPT_NO_SYNC     = 1u << 3, /* with DEVICE_IO: enqueue only, do not synchronise the stream */
    MRAG_OPT_FORCE_MMA128 = 1u << 4, /* pin the 128-query candidate scan + exact rescoring (testing / tuning) */
    MRAG_OPT_COALESCE    = 1u << 5  /* host buffers, no filter, NULL stream: requests from concurrent host threads that arrive
                                       while a scan is in flight are served together by the next scan (one pass over the
                                       corpus for up to 1024 queries).  The reference issues single-query statements from up
                                       to 5 concurrent narrow searches per request plus uvicorn concurrency
                                       (corpus_search_agent.py:794-797); a lone request is served immediately. */
};

/*
 * The scan: for each of nq queries, ORDER BY embedding_vec <=> q LIMIT k over the rows that
 * pass `filter` (NULL = only "embedding_vec IS NOT NULL").
 *   q       nq*dim float32 (query vectors are float4 server side, vector_store.py:272)
 *   scores  nq*k  float32   similarity = 1 - cosine_distance, descending; NaN for zero-norm rows
 *   rows    nq*k  int64     row index + row_base; -1 beyond counts[i]
 *   counts  nq    int32     rows returned (< k when fewer rows pass the filter)
 * Ordering: similarity descending, NaN last, ties by ascending row (pgvector leaves ties
 * unspecified; this library fixes them so results do not depend on sharding).
 * `stream`: cudaStream_t to run on.  NULL means: with host buffers, an internal per-call stream
 * (concurrent callers overlap); with MRAG_OPT_DEVICE_IO, the CUDA default (legacy) stream.
 */
int mrag_search(mrag_index* idx, const float* q, int nq, int k, const mrag_filter* filter,
                float* scores, int64_t* rows, int32_t* counts, uint32_t options, void* stream);

/* --- hybrid rerank (config 5): rerank score fused into the scan --------------------------------
 *
 * Replaces, for ONE arm (vector) over ALL rows that pass the filter, the per-candidate loop of
 * `_rerank` (app/services/corpus_search.py:1909-2297) with `_best_arm_sim` (:1787-1814):
 *
 *   sim'  = max(0, (clamp01(cos) - 0.5) * 2)                                   (:1802-1814, :1569)
 *   cov   = sum_i w_i [phrase_i present] / sum_i w_i                           (:2041-2086)
 *           present = doc carries the phrase's j-code  OR  phrase in body/meta haystack
 *   jpd   = min(1, sum_c qcat_c * ccat_c / sum_c qcat_c)                       (:324-377)
 *   score = (w_sim sim' + w_auth auth + w_len len + w_jpd jpd + w_cov cov) / max_weight   (:2113-2118)
 *   score *= boost  if a d: phrase code is among the chunk's chunk_d_tags      (:2120-2137)
 *   drop if cov < floor unless promoted / contact-exact / d-tag matched        (:2183-2247)
 *
 * Everything that needs TEXT (substring tests, regexes, lengths) is evaluated once per row by the
 * host shim when the row is inserted and shipped here as mrag_chunkfeat; the kernels see bits.
 */
#define MRAG_PHRASE_WORDS 2     /* phrase dictionary of <= 128 phrases (the query bank's required phrases) */
#define MRAG_JPD_CATS 11        /* non-empty categories of _JPD_PATTERNS, corpus_search.py:233-309 */
#define MRAG_JTAG_WORDS 4       /* document j: tag codes < 256 */
#define MRAG_HYB_MAX_PHRASES 16

typedef struct mrag_chunkfeat {
    uint64_t phrase_bits[MRAG_PHRASE_WORDS]; /* bit p: dictionary phrase p occurs in the body or meta haystack (:1844-1906) */
    uint8_t  jpd_hits[MRAG_JPD_CATS];        /* patterns of category c found in the body haystack (:339-349)        */
    uint8_t  flags;                          /* MRAG_CF_* */
    float    length_score;                   /* _length_score(text), :1779-1784                                      */
    uint16_t dtags[4];                       /* codes (>= 1) of the keys of chunk_d_tags, 0 = empty slot, models.py:278-280 */
} mrag_chunkfeat;                            /* 40 bytes */

enum {
    MRAG_CF_SHORT_TEXT    = 1u << 0,  /* body haystack has <= 20 words: _classify_jpd scores hits / sqrt(n) (:336-347) */
    MRAG_CF_CONTACT_VALUE = 1u << 1,  /* _CONTACT_VALUE_RE matches the text (:676-683, :2219-2222)                      */
    MRAG_CF_PROMOTED      = 1u << 2,  /* promoted neighbour / bm25_inherited: exempt from the coverage floor (:2205-2208) */
    MRAG_CF_DTAG_OVERFLOW = 1u << 3   /* chunk_d_tags has more than 4 keys: the rest are in the overflow table (mrag_set_dtag_overflow) */
};

typedef struct mrag_hybrid_query {
    int32_t  n_phrases;                              /* required phrases, 0 = no coverage term and no floor       */
    float    phrase_weight[MRAG_HYB_MAX_PHRASES];    /* selectivity weights (>= 0), :1977-1985                    */
    int16_t  phrase_bit[MRAG_HYB_MAX_PHRASES];       /* dictionary index of the phrase, -1 = occurs nowhere       */
    int16_t  phrase_jbit[MRAG_HYB_MAX_PHRASES];      /* doc j-tag bit that gives binary credit, -1 = none (:2055-2063) */
    uint16_t phrase_dcode[MRAG_HYB_MAX_PHRASES];     /* chunk d-tag code (>= 1) of a d: phrase code, 0 = none (:2124-2133) */
    float    qcat[MRAG_JPD_CATS];                    /* _classify_jpd(query); all zero = no jpd term (:1952-1953) */
    float    auth_score[32];                         /* authority code -> _authority_score (:1773-1776); [31] = NULL / unknown */
    float    w_sim, w_auth, w_len, w_jpd, w_cov;     /* 0.25, 0.10, 0.05, 0.20 | 0, 0.55 | 0  (:2006-2011)        */
    float    boost;                                  /* CHUNK_TAG_BOOST (config.py:129)                            */
    float    floor;                                  /* _TAG_COVERAGE_FLOOR (:604)                                 */
    uint32_t contact_query;                          /* _CONTACT_QUERY_RE matched the query (:1959)                */
    /* restrict this query to rows whose source_type code is in the set (bit 255 = NULL); all zero = no restriction.
     * The per-(arm, source_type) decay of :2258-2285 needs each category's own best, so the shim asks for the top
     * rows of every category separately (one query slot per category) and merges after decaying.                */
    uint64_t source_type_any[MRAG_SMALL_WORDS];
} mrag_hybrid_query;

/* per-row text features of rows [first_row, first_row + n) (HOST array) */
int mrag_set_chunk_features(mrag_index* idx, int64_t first_row, const mrag_chunkfeat* feat, int64_t n);
/* chunk_d_tags keys that do not fit the 4 inline slots of mrag_chunkfeat: n (row, code) pairs, HOST arrays, sorted by
 * row (rows carry MRAG_CF_DTAG_OVERFLOW).  `chunk_d_tags ? :key` (corpus_search.py:1637-1672) matches any key of the JSONB
 * map; this call REPLACES the whole table. */
int mrag_set_dtag_overflow(mrag_index* idx, const uint32_t* rows, const uint16_t* codes, int64_t n);
/* document j: tag sets as bitsets, n_docs * MRAG_JTAG_WORDS u64 (HOST), like mrag_set_doc_tags */
int mrag_set_doc_jtags(mrag_index* idx, int64_t first_doc, const uint64_t* bits, int64_t n_docs);
/* The fused scan: like mrag_search, ordered by rerank score DESC (ties by ascending row) over the rows that
 * pass `filter` and the coverage floor.  hq: nq entries (HOST).  scores = rerank scores; cos_out (may be NULL)
 * = clamp01(similarity) of each returned row.  Host buffers only (options: FORCE_* ignored). k <= MRAG_FUSED_K. */
int mrag_search_hybrid(mrag_index* idx, const float* q, int nq, int k, const mrag_filter* filter,
                       const mrag_hybrid_query* hq, float* scores, float* cos_out, int64_t* rows, int32_t* counts,
                       void* stream);

/* `_rerank` over a candidate LIST -- the RRF output of the bm25 / vector / d-tag arms (corpus_search.py:3519-3622), after
 * the enrichment steps attached neighbour text and inherited document tags.  The host shim turns each candidate's
 * haystacks into bits exactly as it does for stored rows; `sim` is `_best_arm_sim` (:1787-1814: max over the arms of the
 * vector arm's max(0, (cos - 0.5) * 2) and the other arms' raw scores -- this is where the BM25 arm's score enters). */
typedef struct mrag_candidate {
    mrag_chunkfeat feat;   /* bits of the candidate's body (+ neighbour text) and meta haystacks (:1850-1906) */
    float    sim;          /* _best_arm_sim */
    uint32_t doc_idx;      /* its document: binary j-tag credit (:2055-2063) */
    uint8_t  authority;    /* code of authority_level (31+ = unknown) */
    uint8_t  dtag_match;   /* 1 = a d: phrase code is a key of the chunk's chunk_d_tags (:2124-2133) */
    uint8_t  reserved[2];
} mrag_candidate;          /* 52 bytes */
/* cands: n entries (HOST); hq: ONE query.  Outputs (HOST, n each): rerank score before the per-category decay, the
 * weighted coverage, and keep = the coverage floor and its exemptions let the candidate through (:2183-2247). */
int mrag_rerank_candidates(mrag_index* idx, const mrag_candidate* cands, int64_t n, const mrag_hybrid_query* hq,
                           float* scores, float* coverage, uint8_t* keep);

/* The d-tag arm's WHERE (`_dtag_arm`, corpus_search.py:1605-1701): rows that pass `filter` -- evaluated over every
 * live row: that statement has no "embedding_vec IS NOT NULL" -- and whose chunk_d_tags hold any of `dcodes`
 * (n_codes <= 32).  host_mask_out: ceil(size/32) words (HOST); counts (HOST, n_codes + 1): [0] = rows passing the
 * filter (the IDF count's n_total, :1649-1653), [1+i] = of those, rows holding code i.  The shim orders the matches
 * by (authority tier, id) and applies LIMIT k (:1674-1680). */
int mrag_dtag_mask(mrag_index* idx, const mrag_filter* filter, const uint16_t* dcodes, int n_codes,
                   uint32_t* host_mask_out, int64_t* counts);

/* Optional hook for callers that pipeline work across streams: `cuda_event` (a cudaEvent_t, or NULL to clear) is
 * recorded in the search's stream right after the PREPARE phase of every following mrag_search on this index (queries
 * padded, filter mask built -- i.e. just before the scan kernels are launched).  The row-sharded searcher
 * (mobius-rag_b200/sharded.py, search_async) uses it to release the PREVIOUS search's cross-rank exchange kernel on a side
 * stream once the next scan is already queued, so the two launches do not compete at the moment the previous search ends. */
int mrag_set_prepared_event(mrag_index* idx, void* cuda_event);

/* Global row id offset added to every returned row (shard base for row-sharded corpora). */
int mrag_set_row_base(mrag_index* idx, int64_t row_base);
/* Explicit ids for rows [first_row, first_row + n) (HOST array; may be set before those rows are appended, up to the
 * capacity): a search returns ids[row] instead of row + row_base.  Used when ONE process owns several shards of one
 * table (a FastAPI worker holding all 8 GPUs of a box, vector_store.py:181-226 behind one store object): the id is the
 * row's position in the host table, so the k-way merge of the shards' lists (mrag_merge_topk: score DESC, id ASC)
 * breaks ties exactly as an unsharded table would.  Within a shard the ids must increase with the row. */
int mrag_set_row_ids(mrag_index* idx, int64_t first_row, const int64_t* ids, int64_t n);

/* K4: k-way merge of `n_lists` per-shard results into one nq*k result on `device`.
 * This is the step that follows the allgather across row shards.  List l lives at
 * (d_scores_in + l*stride_scores)[nq*k], (d_rows_in + l*stride_rows)[nq*k],
 * (d_counts_in + l*stride_counts)[nq]  (strides in ELEMENTS of each array, so the three arrays
 * of one shard may sit in one packed allgather slot).  DEVICE pointers.  n_lists*k <= 16384. */
int mrag_merge_topk(int device, int n_lists, int nq, int k,
                    const float* d_scores_in, const int64_t* d_rows_in, const int32_t* d_counts_in,
                    int64_t stride_scores, int64_t stride_rows, int64_t stride_counts,
                    float* d_scores_out, int64_t* d_rows_out, int32_t* d_counts_out, void* stream);

/* K4 fused with the exchange (row-sharded search without a collective call): every rank's results travel by peer
 * STORES over NVLink into each peer's gather buffer, a per-(peer, query) flag follows, and the same kernel waits for
 * the peers' flags and merges.  peer_bufs: HOST array of `world` (<= 8) device pointers, peer_bufs[r] = rank r's
 * buffer mapped into this process (symmetric memory), each laid out as
 *     2 gather areas of world * slot_bytes  |  flags u32[world][nq]   (zero-initialised once)
 * with a slot = rows i64[nq*k] | scores f32[nq*k] at scores_off | counts i32[nq] at counts_off.  The caller writes its
 * own results into slot `rank` of area (epoch & 1) of ITS buffer (e.g. with mrag_search, MRAG_OPT_DEVICE_IO) and then
 * calls this with epoch = 1, 2, 3, ... on every rank.  Outputs as mrag_merge_topk. */
int mrag_exchange_merge(int device, int world, int rank, int nq, int k, void* const* peer_bufs,
                        int64_t slot_bytes, int64_t scores_off, int64_t counts_off, uint32_t epoch,
                        float* d_scores_out, int64_t* d_rows_out, int32_t* d_counts_out, void* stream);

/* --- candidate pool on the device ------------------------------------------------------------------
 * `build_candidate_pool` (corpus_search_agent.py:1762-1888) intersects per-tag document sets fetched with one SQL
 * statement per tag (`_doc_ids_with_tag`, :1461-1482) and cascades  L1 J&D&P -> L2 J&D -> L3 AHCA&D -> L4 AHCA.  The
 * per-document tag sets already sit in HBM as bitsets (mrag_set_doc_tags / mrag_set_doc_jtags), so all four levels are
 * ONE kernel over the documents; the chosen level stays on the device as a document bitmap that mrag_search takes
 * through mrag_filter.pool (MRAG_F_DOC_POOL_HANDLE) -- no list of <= 5000 UUIDs per search. */
#define MRAG_POOL_LEVELS 4      /* 0 = L1_JDP, 1 = L2_JD, 2 = L3_AHCA_D, 3 = L4_AHCA */
typedef struct mrag_pool_query {
    uint64_t d_all[MRAG_TAG_WORDS];    /* bits of the d: codes of the query in the document tag sets (ALL must be present) */
    uint64_t p_all[MRAG_TAG_WORDS];    /* bits of the p: codes */
    uint64_t j_all[MRAG_JTAG_WORDS];   /* bits of the j: codes in the document j-tag sets */
    uint64_t ahca[MRAG_JTAG_WORDS];    /* bit of j:regulatory_authority.ahca (:1458) */
    int32_t has_j, has_d, has_p;       /* 1 = the query names codes of that kind AND every one of them is known to the
                                          vocabulary (an unknown code matches no document: leave the kind 0 and the level
                                          comes out empty, as the reference's set intersection does) */
    int32_t has_ahca;                  /* 1 = some document carries the AHCA tag */
} mrag_pool_query;
typedef struct mrag_pool mrag_pool;
/* counts[l] = documents at level l (before any cap); counts[MRAG_POOL_LEVELS] = documents carrying every d: code (the
 * cascade tries L3 only when that set is not empty, :1851).  The document sets are those of the tag tables, as in the
 * reference (`document_tags` rows), whether or not a document currently has chunk rows. */
int mrag_pool_build(mrag_index* idx, const mrag_pool_query* q, mrag_pool** out, int64_t counts[MRAG_POOL_LEVELS + 1]);
/* The level the handle stands for from now on; at most `cap` documents are kept (lowest document indices; the reference
 * keeps list(set)[:5000], :1818); *n_kept (may be NULL) = documents kept. */
int mrag_pool_select(mrag_pool* pool, int level, int64_t cap, int64_t* n_kept);
/* Union extra documents into the selected level (the inherited-authority union, :1966-2002). */
int mrag_pool_add_docs(mrag_pool* pool, const uint32_t* docs, int64_t n);
/* Document indices of the selected level, ascending, at most `max` (HOST out); *n = how many were written. */
int mrag_pool_docs(mrag_pool* pool, uint32_t* out, int64_t max, int64_t* n);
int mrag_pool_destroy(mrag_pool* pool);

/* K2 alone: evaluate `filter` into a row bitmap (bit r of word r/32), DEVICE pointer,
 * ceil(size/32) words; *n_pass (HOST, may be NULL) gets the popcount. */
int mrag_filter_mask(mrag_index* idx, const mrag_filter* filter, uint32_t* d_mask_out,
                     int64_t* n_pass, void* stream);

/* per-kernel device time of the LAST mrag_search on this thread, CUDA events, ms.
 * what: 0 = prepare (query prep + filter-mask), 1 = scan (score + per-producer select),
 *       2 = merge/finalize (+ NaN tail), 3 = whole call on the device.
 * Synchronises on the call's end event. Returns a negative value if there is none. */
float mrag_last_kernel_ms(int what);
/* Profiling ring for benchmarks: after mrag_profile_begin(n) the next n searches ON THIS THREAD
 * record their own event set (no synchronisation is added to the search). mrag_profile_read
 * waits for them and writes up to `max` durations (ms) of phase `what` (as above) in call
 * order; returns how many were written. mrag_profile_begin(0) releases the ring. */
int mrag_profile_begin(int n);
int mrag_profile_read(int what, float* out_ms, int max);
/* number of kernels this library has launched in this process so far */
int64_t mrag_launch_count(void);
/* name of the scan kernel variant the last search on this thread used ("gemv", "mma") */
const char* mrag_last_scan_kind(void);

const char* mrag_last_error(void);
const char* mrag_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MRAG_H_ */
