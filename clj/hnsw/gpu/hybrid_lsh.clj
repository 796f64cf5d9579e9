(ns hnsw.gpu.hybrid-lsh
  "Drop-in for hnsw.ann.hash.hybrid-lsh (src/hnsw/ann/hash/hybrid_lsh.clj): same build-index / search-knn /
   search-hybrid-multiprobe / index-info, same projections (java.util.Random(42).nextGaussian on the JVM itself — the
   Python mirror needs hb_lsh_matrices for that), same hash tables (host maps bucket-id -> row indices in insertion
   order).  The device does the arithmetic: hash bits = hb_pairwise (HB_IP) against the first NUM-HASH-BITS rows of every
   projection matrix, bucket scans = hb_gather_score over the probed buckets' members.  The reference's per-bucket
   (take (* k 2)) never reaches the final k (DESIGN.md §7), so whole buckets are scored and sorted once."
  (:require [hnsw.gpu.ffi :as ffi])
  (:import [java.lang.foreign Arena MemorySegment ValueLayout]
           [java.lang.invoke MethodHandle]
           [java.util Random]))

(def ^:const NUM-HASH-TABLES 8)   ; hybrid_lsh.clj:12
(def ^:const NUM-HASH-BITS 12)    ; :13
(def ^:const PROJECTION-DIM 64)   ; :14

(defrecord GpuHybridIndex [handle ids d proj-seg tables ^Arena arena])

(defn- projection-rows
  "generate-random-matrix x 8 from one Random(42) (:24-31, :77-81); keeps rows 0..11 of every table (the only ones
   hash-to-bucket-id reads, :47-55) as one [96 x d] fp64 segment, but draws all 64 rows to keep the stream aligned."
  ^MemorySegment [^Arena arena d]
  (let [rng (Random. 42)
        seg (.allocate arena (* 8 (long NUM-HASH-TABLES) (long NUM-HASH-BITS) (long d)) 64)]
    (dotimes [t NUM-HASH-TABLES]
      (dotimes [i PROJECTION-DIM]
        (dotimes [j d]
          (let [g (.nextGaussian rng)]
            (when (< i NUM-HASH-BITS)
              (.setAtIndex seg ValueLayout/JAVA_DOUBLE (+ (* (+ (* t NUM-HASH-BITS) i) (long d)) j) g))))))
    seg))

(defn- bucket-ids
  "compute-hash-vector + hash-to-bucket-id (:33-55) for n vectors in `seg` (fp64 [n x d]): vector of [b0 .. b7]."
  [^Arena arena ^MemorySegment seg n d ^MemorySegment proj]
  (let [cols (* NUM-HASH-TABLES NUM-HASH-BITS)
        out (.allocate arena (* 8 (long n) cols) 8)]
    (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-pairwise
                                      (object-array [seg (long n) (int ffi/F64) proj (long cols) (int ffi/F64)
                                                     (int d) (int ffi/IP) out])))
    (vec (for [r (range n)]
           (vec (for [t (range NUM-HASH-TABLES)]
                  (reduce (fn [id i]
                            (if (>= (.getAtIndex out ValueLayout/JAVA_DOUBLE (+ (* r cols) (* t NUM-HASH-BITS) i)) 0.0)
                              (bit-or id (bit-shift-left 1 i))
                              id))
                          0 (range NUM-HASH-BITS))))))))

(defn build-index
  "(build-index data & opts), :66-145, :345-348: data = seq of [id double[]]."
  [data & _opts]
  (let [arena (Arena/ofShared)
        ids (mapv first data)
        d (alength ^doubles (second (first data)))
        rows (ffi/doubles->segment arena (map second data) d)
        proj (projection-rows arena d)
        out (.allocate arena ValueLayout/ADDRESS)]
    (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-flat-create
                                      (object-array [rows (long (count ids)) (int d) (int ffi/F64) (int ffi/COSINE) out])))
    (let [b (bucket-ids arena rows (count ids) d proj)
          tables (vec (for [t (range NUM-HASH-TABLES)]
                        (reduce (fn [m r] (update m (nth (nth b r) t) (fnil conj []) r)) {} (range (count ids)))))]
      (->GpuHybridIndex (.get out ValueLayout/ADDRESS 0) ids d proj tables arena))))

(defn search-hybrid-multiprobe
  "(search-hybrid-multiprobe index query k :num-probes 6 :probe-radius 2), :261-342."
  [^GpuHybridIndex index ^doubles query k & {:keys [num-probes probe-radius] :or {num-probes 6 probe-radius 2}}]
  (with-open [arena (Arena/ofConfined)]
    (let [d (:d index)
          q (ffi/doubles->segment arena [query] d)
          qb (first (bucket-ids arena q 1 d (:proj-seg index)))
          mask (dec (bit-shift-left 1 NUM-HASH-BITS))
          cand (distinct                                   ; first occurrence kept, :327-332
                (for [t (range (min num-probes NUM-HASH-TABLES))
                      b (cons (nth qb t)
                              (for [bit (range (min probe-radius NUM-HASH-BITS))]
                                (bit-and (bit-xor (nth qb t) (bit-shift-left 1 bit)) mask)))
                      r (get (nth (:tables index) t) b [])]
                  r))
          n (count cand)]
      (if (zero? n)
        []
        (let [pq (.allocate arena (* 4 (long n)) 4)
              pr (.allocate arena (* 4 (long n)) 4)
              out (.allocate arena (* 8 (long n)) 8)]
          (doseq [[i r] (map-indexed vector cand)]
            (.setAtIndex pq ValueLayout/JAVA_INT (long i) (int 0))
            (.setAtIndex pr ValueLayout/JAVA_INT (long i) (int r)))
          (ffi/check! (.invokeWithArguments ^MethodHandle ffi/hb-gather-score
                                            (object-array [(:handle index) q (int ffi/F64) (long 1) pq pr (long n) out])))
          (->> (map-indexed (fn [i r] {:id (nth (:ids index) r)
                                       :distance (.getAtIndex out ValueLayout/JAVA_DOUBLE (long i))})
                            cand)
               (sort-by :distance)                         ; stable, like Collections/sort at :334-338
               (take k)
               vec))))))

(defn search-knn
  "(search-knn index query k) / (search-knn index query k mode), :350-364."
  ([index query k] (search-hybrid-multiprobe index query k :num-probes 6 :probe-radius 2))
  ([index query k mode]
   (let [[p r] (case mode :turbo [2 1] :fast [4 1] :balanced [6 2] :accurate [8 3] :precise [8 4] [6 2])]
     (search-hybrid-multiprobe index query k :num-probes p :probe-radius r))))

(defn index-info [^GpuHybridIndex index]
  (let [total (reduce + (map count (:tables index)))]
    {:type "Hybrid LSH Index" :vectors (count (:ids index)) :hash-tables NUM-HASH-TABLES
     :buckets-per-table (bit-shift-left 1 NUM-HASH-BITS) :total-buckets total
     :avg-bucket-size (if (pos? total) (/ (count (:ids index)) total) 0)}))

(defn close! [^GpuHybridIndex index]
  (.invokeWithArguments ^MethodHandle ffi/hb-index-free (object-array [(:handle index)]))
  (.close ^Arena (:arena index)))
