"""Mirror of hnsw.helper.index-io (src/hnsw/helper/index_io.clj:10-80) and hnsw.api save-index / load-index
(src/hnsw/api.clj:40-50, unimplemented there) for the device-resident index types.

The reference writes the HNSW graph as EDN text (every double through pr-str; 492.9 MB for 31 k vectors,
README.md:22) and cannot persist IVF-FLAT at all.  Here `filepath` holds the device layout verbatim
(hb_index_save: header + tagged sections) and `filepath + '.ids.json'` the String ids the host shim owns, so a load is
file -> pinned staging -> HBM: no k-means, no norm pass, and the loaded index answers with the same bits."""
from __future__ import annotations

import ctypes as C
import json
import os

from . import _lib as hb
from .flat import FlatIndex
from .index import DeviceIndex, new_handle
from .ivf_flat import IVFFlatIndex
from .ultra_fast import HnswIndex

_METRIC_NAMES = {hb.COSINE: "cosine", hb.L2: "euclidean", hb.IP: "ip"}


def save_index(index: DeviceIndex, filepath: str) -> DeviceIndex:
    """(save-index index filepath) -> index (index_io.clj:10-39)."""
    if getattr(index, "_h", None) is None:
        raise hb.HbInvalid(hb.ERR_INVALID, "index is closed")
    hb.check(hb.lib().hb_index_save(index._h, os.fsencode(filepath)))
    side = filepath + ".ids.json"
    if index.ids is not None:
        with open(side + ".tmp", "w") as f:
            json.dump(index.ids, f)
        os.replace(side + ".tmp", side)
    elif os.path.exists(side):
        os.remove(side)
    # the index flavour is a property of the host mirror (a Lightning index is an IVF-FLAT layout probed by other tables)
    meta = filepath + ".meta.json"
    with open(meta + ".tmp", "w") as f:
        json.dump({"class": type(index).__name__, "distance_fn": getattr(index, "distance_fn", None)}, f)
    os.replace(meta + ".tmp", meta)
    return index


def load_index(filepath: str, distance_fn=None):
    """(load-index filepath distance-fn) (index_io.clj:41-80): a missing file returns None like the reference's
    nil (:78-80); a truncated or foreign file raises.  `distance_fn` is accepted for signature parity — the metric
    is part of the file."""
    if not os.path.exists(filepath):
        return None
    h = new_handle()
    hb.check(hb.lib().hb_index_load(os.fsencode(filepath), C.byref(h)))
    ids = None
    side = filepath + ".ids.json"
    if os.path.exists(side):
        with open(side) as f:
            ids = json.load(f)
    base = DeviceIndex(h.value, ids)
    info = base.info()
    metric = _METRIC_NAMES[info["metric"]]
    cls = {hb.INDEX_FLAT: FlatIndex, hb.INDEX_IVF_FLAT: IVFFlatIndex, hb.INDEX_HNSW: HnswIndex}[info["type"]]
    meta = filepath + ".meta.json"
    if os.path.exists(meta):
        with open(meta) as f:
            m = json.load(f)
        if cls is IVFFlatIndex and m.get("class") == "LightningIndex":
            from .lightning import LightningIndex

            cls = LightningIndex
        if m.get("distance_fn"):
            metric = m["distance_fn"]
    ix = cls.__new__(cls)
    DeviceIndex.__init__(ix, h.value, ids)
    base._h = None  # ownership moved to ix
    if cls is FlatIndex:
        ix.metric = info["metric"]
    elif issubclass(cls, IVFFlatIndex):
        ix.num_partitions = info["nlist"]
        ix.distance_fn = metric
    return ix
