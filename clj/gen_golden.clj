(ns gen-golden
  "Emits tests/golden/reference_golden.json from the REFERENCE itself (damesek/hnsw-clj), for whoever holds a JVM:

     cd <hnsw-clj checkout>
     clojure -Sdeps '{:paths [\"src\" \"test\" \"<this repo>/clj\"]}' -M -m gen-golden <this repo>/tests/golden/reference_golden.json

   The build image of this repo has no JVM, so the committed fixtures (tests/golden/golden.json) come from the CPU oracle and
   parity with the reference is pinned only by its three pairwise known-answer tests.  This script closes that gap: it runs
   the reference's own build-index / search-knn / compute-exact-knn / kmeans-plus-plus-init on the same seeded inputs as
   tests/golden/make_golden.py (data-generator/generate-dataset, seeds 42 / 43, values rounded to float like the double[]
   of fp32 embeddings the reference is fed) and writes the same keys.  tests/test_golden.py::test_reference_goldens_when_present
   then checks the oracle — and through it the CUDA path — against the reference's numbers, bit for bit.
   Doubles are written with Double/toHexString (Python: float.fromhex)."
  (:require [clojure.data.json :as json]
            [data-generator :as gen]
            [hnsw.ann.partition.ivf-flat :as ivf]
            [hnsw.bench :as bench]
            [hnsw.simd-optimized :as simd-opt]
            [hnsw.ultra-fast :as ultra])
  (:import [hnsw.ann.partition.ivf_flat IVFFlatIndex]))

(defn- hex [^double x] (Double/toHexString x))

(defn- as-f32
  "[id double[]] pairs whose values are fp32-representable, like embeddings exported as float32"
  [data]
  (mapv (fn [[id ^doubles v]] [id (double-array (map #(double (float %)) v))]) data))

(defn- dataset [n d kind seed clusters]
  (as-f32 (gen/generate-dataset n d :distribution kind :num-clusters clusters :noise-level 0.1 :seed seed :format :indexed)))

(defn- row-of [id] (Long/parseLong (subs id 4)))  ; "vec_17" -> 17

(defn- top-k-by
  "all distances, stable sort, take k (the shape of compute-exact-knn, src/hnsw/bench.clj:72-84) with another distance"
  [data ^doubles q k f]
  (vec (take k (sort-by :distance (map (fn [[id v]] {:id id :distance (f q v)}) data)))))

(defn- flat-cases []
  (let [data (dataset 500 32 :unit 42 10)
        queries (mapv second (dataset 12 32 :unit 43 10))
        emit (fn [results] {:ids (mapv (fn [r] (mapv #(row-of (:id %)) r)) results)
                            :dist (vec (mapcat (fn [r] (map #(hex (:distance %)) r)) results))})]
    {"flat_unit/cosine" (emit (mapv #(bench/compute-exact-knn data % 10) queries))
     "flat_unit/euclidean" (emit (mapv #(top-k-by data % 10 ultra/euclidean-distance-ultra) queries))
     ;; inner product ranks by descending dot: distance = -dot (an extension of this repo, BASELINE configs[2])
     "flat_unit/ip" (emit (mapv #(top-k-by data % 10 (fn [a b] (- (double (simd-opt/dot-product a b))))) queries))
     "flat_unit/row0" (mapv hex (second (first data)))}))

(defn- ivf-case []
  (let [data (dataset 1200 24 :clustered 42 12)
        queries (mapv second (dataset 16 24 :clustered 43 12))
        dist-fn ultra/cosine-distance-ultra
        seeds (#'ivf/kmeans-plus-plus-init data 16 dist-fn)
        seed-rows (mapv (fn [c] (first (keep-indexed (fn [i [_ v]] (when (identical? v c) i)) data))) seeds)
        ^IVFFlatIndex index (ivf/build-index data :num-partitions 16 :max-iterations 10 :show-progress? false)
        parts (.partitions index)
        cents (.centroids index)
        assign (let [a (long-array (count data))]
                 (doseq [[p members] (map-indexed vector parts) [id _] members] (aset a (row-of id) (long p)))
                 (vec a))
        results (mapv #(ivf/search-knn index % 10 :balanced) queries)   ; :balanced = 4 probes (ivf_flat.clj:243-247)
        probes (mapv (fn [q] (->> cents
                                  (map-indexed (fn [i c] {:idx i :dist (dist-fn q c)}))
                                  (sort-by :dist) (take 4) (mapv :idx)))   ; the selection at ivf_flat.clj:261-269
                     queries)
        exact (mapv #(bench/compute-exact-knn data % 10) queries)
        recall (/ (reduce + (map (fn [a e] (double (bench/calc-recall a e))) results exact)) (count queries))]
    {"ivf_clustered" {:seed_rows seed-rows
                      :assign assign
                      :centroid0 (mapv hex (first cents))
                      :centroid_sum (hex (reduce + (mapcat seq cents)))
                      :ids (mapv (fn [r] (mapv #(row-of (:id %)) r)) results)
                      :dist (vec (mapcat (fn [r] (map #(hex (:distance %)) r)) results))
                      :probes probes
                      :recall recall}}))

(defn -main [& [out]]
  (let [golden (merge (flat-cases) (ivf-case)
                      {"pairwise" {:cos_123_456 (hex (ultra/cosine-distance-ultra (double-array [1 2 3]) (double-array [4 5 6])))
                                   :euclid_123_456 (hex (ultra/euclidean-distance-ultra (double-array [1 2 3]) (double-array [4 5 6])))}
                       "meta" {:generator "clj/gen_golden.clj run against the reference"
                               :java (System/getProperty "java.version")}})]
    (spit (or out "reference_golden.json") (json/write-str golden))
    (println "wrote" (or out "reference_golden.json"))
    (shutdown-agents)))
