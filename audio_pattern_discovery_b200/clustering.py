"""Mirror of the reference's src/clustering.rs surface on top of apd_upgma (include/apd.h):
a result-identical fast form of the naive UPGMA that consumes the distance matrix
(src/main.rs:196-203).  Host code -- the north_star keeps clustering on the host.

    operations, clusters = AgglomerativeClustering.clustering(distances, n, perc)
    grouped = AgglomerativeClustering.cluster_sets(operations, clusters, n)
"""
import ctypes as C
import enum

import numpy as np

from . import _capi


class Merge(enum.Enum):
    """src/clustering.rs:7-13"""
    Sequence2Sequence = 0
    Sequence2Cluster = 1
    Cluster2Sequence = 2
    Cluster2Cluster = 3


class ClusteringOperation:
    """src/clustering.rs:18-25"""

    def __init__(self, merge_i, merge_j, into, distance, operation, tie=False):
        self.merge_i, self.merge_j, self.into = int(merge_i), int(merge_j), int(into)
        self.distance = np.float32(distance)
        self.operation = operation
        self.tie = bool(tie)  # another root pair had exactly this linkage (HashSet-order dependent upstream)

    def __repr__(self):
        return ("ClusteringOperation { merge_i: %d, merge_j: %d, into: %d, distance: %r, operation: %s }"
                % (self.merge_i, self.merge_j, self.into, float(self.distance), self.operation.name))


class AgglomerativeClustering:
    @staticmethod
    def clustering(distances, n_instances, perc, threshold=None):
        """src/clustering.rs:81-110.  distances: flat n*n float32 (row-major, result[x*n+y]).
        threshold: optional precomputed numerics::percentile(distances, perc) (e.g. from
        Context.percentile on the device).  Returns (operations, set of root cluster ids)."""
        d = np.ascontiguousarray(distances, dtype=np.float32).ravel()
        n = int(n_instances)
        if d.size != n * n:
            raise ValueError("distances must hold n_instances^2 entries")
        ops = (_capi.apd_merge * max(n, 1))()
        n_ops = C.c_uint32(0)
        thr_out = C.c_float(0)
        assign = np.zeros(max(n, 1), dtype=np.uint32)
        thr_in = None
        if threshold is not None:
            thr_in = C.byref(C.c_float(float(threshold)))
        print("\tset parents to self")          # the reference's progress lines (src/clustering.rs:87-102)
        print("\tbuild initial dendrogram")
        print("\testimate threshold")
        st = _capi.lib().apd_upgma(d.ctypes.data_as(C.POINTER(C.c_float)), n, float(perc),
                                   C.cast(thr_in, C.POINTER(C.c_float)) if thr_in is not None else None,
                                   ops, C.byref(n_ops), C.byref(thr_out),
                                   assign.ctypes.data_as(C.POINTER(C.c_uint32)))
        if st != _capi.APD_OK:
            raise IndexError("index out of bounds: percentile(distances, %r) (the reference panics here)" % perc)
        print("Clustering with %s" % thr_out.value)
        operations = [ClusteringOperation(o.merge_i, o.merge_j, o.into, o.distance, Merge(o.operation), o.tie)
                      for o in ops[:n_ops.value]]
        AgglomerativeClustering.last_threshold = np.float32(thr_out.value)
        return operations, set(int(r) for r in assign[:n])

    @staticmethod
    def cluster_sets(operations, cluster_ids, n_instances):
        """src/clustering.rs:40-76 (singular clusters are reported and skipped, like the reference)."""
        results = {}
        for op in operations:
            cluster = list(results.get(op.merge_i, [op.merge_i])) + list(results.get(op.merge_j, [op.merge_j]))
            results[op.into] = cluster
        grouped = []
        for cluster in cluster_ids:
            if cluster in results:
                grouped.append([i for i in results[cluster] if i < n_instances])
            else:
                print("Cluster not found: %d | Singular cluster" % cluster)
        return grouped
