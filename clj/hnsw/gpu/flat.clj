(ns hnsw.gpu.flat
  "Exact flat search on the device: drop-in for hnsw.bench/compute-exact-knn (src/hnsw/bench.clj:72-84) and
   hnsw.simd-optimized/top-k-distances (src/hnsw/simd_optimized.clj:271-280)."
  (:require [hnsw.gpu.ffi :as ffi])
  (:import [java.lang.foreign Arena ValueLayout]
           [java.lang.invoke MethodHandle]))

(defn compute-exact-knn
  "(compute-exact-knn vectors query k): vectors = seq of [id double[]] -> [{:id :distance} ...]"
  [vectors ^doubles query k]
  (with-open [arena (Arena/ofConfined)]
    (let [ids (mapv first vectors)
          d (alength query)
          rows (ffi/doubles->segment arena (map second vectors) d)
          q (ffi/doubles->segment arena [query] d)
          out (.allocate arena ValueLayout/ADDRESS)
          out-ids (.allocate arena (* 8 (long k)) 8)
          out-d (.allocate arena (* 8 (long k)) 8)]
      (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-flat-create
                                        (object-array [rows (long (count ids)) (int d) (int ffi/F64) (int ffi/COSINE) out])))
      (let [h (.get out ValueLayout/ADDRESS 0)]
        (try
          (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-search
                                            (object-array [h q (int ffi/F64) (long 1) (int k) (int 0) out-ids out-d])))
          (vec (for [j (range k)
                     :let [row (.getAtIndex out-ids ValueLayout/JAVA_LONG (long j))]
                     :when (>= row 0)]
                 {:id (nth ids row) :distance (.getAtIndex out-d ValueLayout/JAVA_DOUBLE (long j))}))
          (finally
            (.invokeWithArguments ^MethodHandle ffi/hb-index-free (object-array [h]))))))))
