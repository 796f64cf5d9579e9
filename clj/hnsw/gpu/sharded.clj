(ns hnsw.gpu.sharded
  "Row-sharded multi-GPU index: the scale-out model of hnsw.ann.partition.partitioned-hnsw
   (src/hnsw/ann/partition/partitioned_hnsw.clj:86-196 — independent sub-indexes over row ranges, searched in parallel,
   results concatenated, sort-by :distance, take k) applied to ONE global IVF-FLAT or flat index.

   Deployment: one JVM per GPU of the box (rank g of G).  Rank 0 draws the communicator id (hb_comm_unique_id, 128 bytes)
   and hands it to the others by any means (a file, an env var, a socket); every rank then calls (init! ...) once.
   (build-index ...) and (search-batch ...) are collective: every rank makes the same call with its own rows / the same
   queries and gets the same global result.  The exchange of the local top-k lists and the merge run inside the library
   (NVLink peer windows or ncclAllGather + merge kernel); nothing crosses the JVM boundary but the final [nq x k] arrays."
  (:require [hnsw.gpu.ffi :as ffi])
  (:import [java.lang.foreign Arena MemorySegment ValueLayout]
           [java.lang.invoke MethodHandle]
           [java.nio.file Files Paths]))

(defn unique-id
  "Rank 0: the 128-byte id to share with the other ranks."
  ^bytes []
  (with-open [arena (Arena/ofConfined)]
    (let [seg (.allocate arena 128 8)]
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-comm-unique-id (object-array [seg])))
      (.toArray seg ValueLayout/JAVA_BYTE))))

(defn init!
  "After (ffi hb_init device): joins the communicator.  id = (unique-id) of rank 0."
  [^bytes id nranks rank]
  (with-open [arena (Arena/ofConfined)]
    (let [seg (.allocate arena 128 8)]
      (MemorySegment/copy id 0 seg ValueLayout/JAVA_BYTE 0 128)
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-comm-init (object-array [seg (int nranks) (int rank)]))))))

(defn init-from-file!
  "Convenience: rank 0 writes the id to `path`, the others wait for it."
  [path nranks rank]
  (let [p (Paths/get path (make-array String 0))]
    (if (zero? rank)
      (let [id (unique-id)] (Files/write p id (make-array java.nio.file.OpenOption 0)) (init! id nranks rank))
      (do (while (or (not (Files/exists p (make-array java.nio.file.LinkOption 0))) (< (Files/size p) 128)) (Thread/sleep 10))
          (init! (Files/readAllBytes p) nranks rank)))))

(defrecord ShardedIndex [^MemorySegment handle ids first-row dim])

(defn build-index
  "This rank's rows `data` = seq of [id double[]] for the global rows [first-row, first-row + (count data)).
   seed-rows = the k-means++ result (global row indices, identical on every rank)."
  [data first-row & {:keys [num-partitions max-iterations seed-rows] :or {num-partitions 24 max-iterations 10}}]
  (with-open [arena (Arena/ofConfined)]
    (let [d (alength ^doubles (second (first data)))
          rows (ffi/doubles->segment arena (map second data) d)
          seeds (.allocate arena (* 8 (long num-partitions)) 8)
          out (.allocate arena 8 8)]
      (doseq [[i s] (map-indexed vector seed-rows)] (.setAtIndex seeds ValueLayout/JAVA_LONG (long i) (long s)))
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-sharded-ivf-build
                                        (object-array [rows (long (count data)) (int d) (int ffi/F64) (int ffi/COSINE)
                                                       (int num-partitions) (int max-iterations) seeds (long first-row) out])))
      (->ShardedIndex (.get out ValueLayout/ADDRESS 0) (mapv first data) first-row d))))

(defn search-batch
  "Global top-k for every query: [[{:row <global row> :distance d} ...] ...], identical on every rank.  The caller maps
   global rows back to ids with its own (first-row, ids) tables."
  [^ShardedIndex index queries k num-probes]
  (with-open [arena (Arena/ofConfined)]
    (let [nq (count queries)
          q (ffi/doubles->segment arena queries (:dim index))
          ids (.allocate arena (* 8 (long nq) (long k)) 8)
          dist (.allocate arena (* 8 (long nq) (long k)) 8)]
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-sharded-search
                                        (object-array [(:handle index) q (int ffi/F64) (long nq) (int k) (int num-probes) ids dist])))
      (vec (for [i (range nq)]
             (vec (for [j (range k)
                        :let [row (.getAtIndex ids ValueLayout/JAVA_LONG (+ (* (long i) (long k)) (long j)))]
                        :when (>= row 0)]
                    {:row row :distance (.getAtIndex dist ValueLayout/JAVA_DOUBLE (+ (* (long i) (long k)) (long j)))})))))))
