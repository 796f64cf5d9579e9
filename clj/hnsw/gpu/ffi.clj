(ns hnsw.gpu.ffi
  "java.lang.foreign (JDK >= 22) downcall handles for libhnswb200.so — the C ABI declared in include/hnswb200.h.
   One handle per entry point; every compute call returns an int32 status (0 = OK) and hb_last_error gives the
   message.  No CPU fallback: a missing library or device surfaces as ex-info."
  (:import [java.lang.foreign Arena FunctionDescriptor Linker MemorySegment SymbolLookup ValueLayout]
           [java.lang.invoke MethodHandle]))

(set! *warn-on-reflection* true)

(def ^:const F32 0) (def ^:const BF16 1) (def ^:const F64 2)
(def ^:const COSINE 0) (def ^:const L2 1) (def ^:const IP 2)

(defonce ^Linker linker (Linker/nativeLinker))
(defonce ^Arena lib-arena (Arena/global))
(defonce ^SymbolLookup lookup
  (SymbolLookup/libraryLookup (or (System/getProperty "hnswb200.lib") "libhnswb200.so") lib-arena))

(def ^:private I ValueLayout/JAVA_INT)
(def ^:private J ValueLayout/JAVA_LONG)
(def ^:private P ValueLayout/ADDRESS)

(defn- handle ^MethodHandle [^String sym ret & args]
  (let [addr (.orElseThrow (.find lookup sym))
        fd (if ret
             (FunctionDescriptor/of ret (into-array java.lang.foreign.MemoryLayout args))
             (FunctionDescriptor/ofVoid (into-array java.lang.foreign.MemoryLayout args)))]
    (.downcallHandle linker addr fd (make-array java.lang.foreign.Linker$Option 0))))

;; int hb_init(int device); const char* hb_last_error(void)
(defonce hb-init (handle "hb_init" I I))
(defonce hb-last-error (handle "hb_last_error" P))
;; int hb_row_norms(const void* rows, int64 n, int32 d, int dtype, double* out)
(defonce hb-row-norms (handle "hb_row_norms" I P J I I P))
;; int hb_pairwise(a, na, adtype, b, nb, bdtype, d, metric, out)
(defonce hb-pairwise (handle "hb_pairwise" I P J I P J I I I P))
;; int hb_flat_create(rows, n, d, dtype, metric, hb_index** out)
(defonce hb-flat-create (handle "hb_flat_create" I P J I I I P))
;; int hb_ivf_build(rows, n, d, dtype, metric, nlist, iters, seed, hb_index** out)
(defonce hb-ivf-build (handle "hb_ivf_build" I P J I I I I I J P))
;; int hb_lightning_build(rows, n, d, dtype, metric, nlist, seed, hb_index** out)
(defonce hb-lightning-build (handle "hb_lightning_build" I P J I I I I J P))
;; int hb_search(index, queries, qdtype, nq, k, param, int64* out_ids, double* out_dist)
(defonce hb-search (handle "hb_search" I P P I J I I P P))
;; int hb_gather_score(index, queries, qdtype, nq, pair_query, pair_row, npairs, out)
(defonce hb-gather-score (handle "hb_gather_score" I P P I J P P J P))
;; int hb_index_save(const hb_index*, const char* path); int hb_index_load(const char* path, hb_index** out)
(defonce hb-index-save (handle "hb_index_save" I P P))
(defonce hb-index-load (handle "hb_index_load" I P P))
;; int hb_index_free(hb_index*)
(defonce hb-index-free (handle "hb_index_free" I P))
;; int hb_index_set_mode(hb_index*, int mode): HB_MODE_EXACT 0 / HB_MODE_FAST 1 / -1 = process default
(defonce hb-index-set-mode (handle "hb_index_set_mode" I P I))

;; ---- multi-GPU: one JVM per GPU (include/hnswb200.h, "multi-GPU") ----
;; int hb_comm_unique_id(void* out128); int hb_comm_init(const void* id128, int32 nranks, int32 rank)
(defonce hb-comm-unique-id (handle "hb_comm_unique_id" I P))
(defonce hb-comm-init (handle "hb_comm_init" I P I I))
(defonce hb-comm-shutdown (handle "hb_comm_shutdown" I))
;; int hb_comm_broadcast(void* buf, int64 bytes, int32 root)
(defonce hb-comm-broadcast (handle "hb_comm_broadcast" I P J I))
;; int hb_sharded_ivf_build(rows, n_local, d, dtype, metric, nlist, iters, const int64* seed_rows, int64 first_global_row, hb_index** out)
(defonce hb-sharded-ivf-build (handle "hb_sharded_ivf_build" I P J I I I I I P J P))
;; int hb_index_set_id_base(hb_index*, int64 first_global_row)
(defonce hb-index-set-id-base (handle "hb_index_set_id_base" I P J))
;; int hb_sharded_search(index, queries, qdtype, nq, k, param, int64* out_ids, double* out_dist): ids are global rows
(defonce hb-sharded-search (handle "hb_sharded_search" I P P I J I I P P))

;; ---- float[] Vector-API variants + PCAF (src/hnsw/simd.clj:18-115, src/hnsw/ann/dimreduct/pcaf.clj) ----
;; int hb_pairwise_f32lanes(const float* a, int64 na, const float* b, int64 nb, int32 d, int metric, int32 lanes, double* out)
(defonce hb-pairwise-f32lanes (handle "hb_pairwise_f32lanes" I P J P J I I I P))
;; int hb_pcaf_matrix(int32 original_dim, int32 target_dim, int64 seed, float* out)
(defonce hb-pcaf-matrix (handle "hb_pcaf_matrix" I I I J P))
;; int hb_pcaf_project(matrix, original_dim, target_dim, rows, n, lanes, float* out)
(defonce hb-pcaf-project (handle "hb_pcaf_project" I P I I P J I P))
;; int hb_pcaf_search(high, low, queries, low_queries, nq, k, k_filter, lanes, int64* out_ids, double* out_dist)
(defonce hb-pcaf-search (handle "hb_pcaf_search" I P P P P J I I I P P))

(defn last-error ^String []
  (let [^MemorySegment p (.invokeWithArguments ^MethodHandle hb-last-error (object-array 0))]
    (.getString (.reinterpret p 4096) 0)))

(defn check! [status]
  (when-not (zero? (int status))
    (if (= -1 (int status))
      (throw (IllegalArgumentException. (last-error)))          ; src/hnsw/api/simple.clj:13-14
      (throw (ex-info (last-error) {:status status})))))

(defn doubles->segment
  "Copies a seq of double[] (the reference's vector representation, src/hnsw/ultra_fast.clj:100) into one
   contiguous fp64 [n x d] segment.  Pass dtype F64; use floats->segment when the values are known fp32."
  ^MemorySegment [^Arena arena vectors d]
  (let [n (count vectors)
        seg (.allocate arena (* 8 (long n) (long d)) 64)]
    (doseq [[i ^doubles v] (map-indexed vector vectors)]
      (MemorySegment/copy v 0 seg ValueLayout/JAVA_DOUBLE (* 8 (long i) (long d)) (int d)))
    seg))
