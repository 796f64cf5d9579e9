(ns hnsw.gpu.ivf-flat
  "Drop-in for hnsw.ann.partition.ivf-flat (src/hnsw/ann/partition/ivf_flat.clj:300-327): same build-index /
   search-knn / index-info signatures and option keys, device-resident index behind libhnswb200.so.
   Adds search-batch (BatchSearchIndex/search-batch*, src/hnsw/api/protocol.clj:58-67): ONE device call for all queries."
  (:require [clojure.edn]
            [hnsw.gpu.ffi :as ffi]
            [hnsw.api.protocol :as proto])
  (:import [java.lang.foreign Arena MemorySegment ValueLayout]
           [java.lang.invoke MethodHandle]))

(defrecord GpuIVFFlatIndex [^MemorySegment handle ids dim num-partitions distance-fn])

(def ^:private mode->probes {:turbo 1 :fast 2 :balanced 4 :accurate 8 :precise 12}) ; ivf_flat.clj:243-247

(defn- metric-code [distance-fn]
  (case distance-fn (:cosine nil) ffi/COSINE :euclidean ffi/L2
        (throw (IllegalArgumentException. (str "unknown :distance-fn " distance-fn)))))

(defn build-index
  "data = seq of [id double[]] pairs (test/data_generator.clj:84-87).  Options as in the reference
   (:num-partitions 24, :max-iterations 10, :distance-fn, ivf_flat.clj:139-148); :distance-fn is a keyword here."
  [data & {:keys [num-partitions distance-fn max-iterations seed]
           :or {num-partitions 24 max-iterations 10 seed 42}}]
  (when (empty? data) (throw (IllegalArgumentException. "cannot build an IVF-FLAT index from no vectors")))
  (with-open [arena (Arena/ofConfined)]
    (let [ids (mapv first data)
          d (alength ^doubles (second (first data)))
          rows (ffi/doubles->segment arena (map second data) d)
          out (.allocate arena ValueLayout/ADDRESS)]
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-ivf-build
                                        (object-array [rows (long (count ids)) (int d) (int ffi/F64)
                                                       (int (metric-code distance-fn)) (int num-partitions)
                                                       (int max-iterations) (long seed) out])))
      (->GpuIVFFlatIndex (.get out ValueLayout/ADDRESS 0) ids d num-partitions distance-fn))))

(defn- search* [^GpuIVFFlatIndex index queries k nprobe]
  (with-open [arena (Arena/ofConfined)]
    (let [nq (count queries)
          q (ffi/doubles->segment arena queries (:dim index))
          out-ids (.allocate arena (* 8 (long nq) (long k)) 8)
          out-d (.allocate arena (* 8 (long nq) (long k)) 8)]
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-search
                                        (object-array [(:handle index) q (int ffi/F64) (long nq) (int k) (int nprobe)
                                                       out-ids out-d])))
      (vec (for [qi (range nq)]
             (vec (for [j (range k)
                        :let [row (.getAtIndex out-ids ValueLayout/JAVA_LONG (+ (* (long qi) k) j))]
                        :when (>= row 0)]                       ; k > n -> n results, test/hnsw/core_test.clj:90-96
                    {:id (nth (:ids index) row)
                     :distance (.getAtIndex out-d ValueLayout/JAVA_DOUBLE (+ (* (long qi) k) j))})))))))

(defn- probes-for [mode num-probes] (or (mode->probes mode) num-probes 4))   ; ivf_flat.clj:249-251

(defn search-knn
  ([index query k] (search-knn index query k :balanced))
  ([index ^doubles query k mode & {:keys [num-probes]}]
   (first (search* index [query] k (probes-for mode num-probes)))))

(defn search-batch
  ([index queries k] (search-batch index queries k :balanced))
  ([index queries k mode & {:keys [num-probes]}]
   (search* index queries k (probes-for mode num-probes))))

(defn index-info [^GpuIVFFlatIndex index]
  {:type "IVF-FLAT (B200)" :vectors (count (:ids index)) :partitions (:num-partitions index)})

(defn save-index
  "Device layout -> `filepath` (hb_index_save) + the String ids as EDN beside it.  The reference has no IVF-FLAT
   persistence; the signature follows hnsw.helper.index-io/save-index (src/hnsw/helper/index_io.clj:10-39)."
  [^GpuIVFFlatIndex index ^String filepath]
  (with-open [arena (Arena/ofConfined)]
    (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-index-save
                                      (object-array [(:handle index) (.allocateFrom arena filepath)]))))
  (spit (str filepath ".ids.edn") (pr-str {:ids (:ids index) :dim (:dim index) :num-partitions (:num-partitions index)
                                           :distance-fn (:distance-fn index)}))
  index)

(defn load-index
  "nil when the file does not exist, like hnsw.helper.index-io/load-index (src/hnsw/helper/index_io.clj:78-80)."
  [^String filepath]
  (when (.exists (java.io.File. filepath))
    (with-open [arena (Arena/ofConfined)]
      (let [out (.allocate arena ValueLayout/ADDRESS)
            meta (clojure.edn/read-string (slurp (str filepath ".ids.edn")))]
        (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-index-load
                                          (object-array [(.allocateFrom arena filepath) out])))
        (->GpuIVFFlatIndex (.get out ValueLayout/ADDRESS 0) (:ids meta) (:dim meta) (:num-partitions meta)
                           (:distance-fn meta))))))

(defn close! [^GpuIVFFlatIndex index]
  (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-index-free (object-array [(:handle index)]))))

(extend-type GpuIVFFlatIndex
  proto/ANNIndex
  (search-knn* [this query k mode] (search-knn this query k (or mode :balanced)))
  (index-info* [this] (index-info this))
  (index-type* [_] :gpu-ivf-flat)
  proto/BatchSearchIndex
  (search-batch* [this queries k mode] (search-batch this queries k (or mode :balanced))))
