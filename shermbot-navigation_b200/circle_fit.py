"""Host-side mirror of the reference's ``circle_fit`` library API over the C ABI (include/nuslam_b200.h).

Keeps the function names of ``nuslam/include/nuslam/circle_fit_library.hpp:18-28`` -- ``clusterPoints``, ``classifyCluster``,
``circleFit`` -- plus the batched ``scan_detect`` that replays ``Landmarks::main_loop`` (nuslam/src/landmarks.cpp:84-109) for S
scans in one launch. Arguments are numpy arrays (host) or torch CUDA tensors (device pointers, no copies).

Every function runs hand-written sm_100a kernels through ``libnuslam_b200.so``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .nuslam import NUSLAM_DEVICE, NUSLAM_HOST, SCAN_UB, NuslamError, _check, _is_torch, lib

BEAMS = 360
FIT_MOMENT, FIT_JACOBI = 0, 1   # NUSLAM_FIT_* of include/nuslam_b200.h


def set_fit(mode) -> int:
    """Arithmetic of circleFit on the batched paths: ``"moment"`` (default: Hyper fit from warp-shuffle moment reductions, circles
    within 1e-9 of the reference) or ``"jacobi"`` (SVD -> eig_sym -> solve in the oracle's operation order). Returns the previous mode."""
    code = {"moment": FIT_MOMENT, "jacobi": FIT_JACOBI}.get(mode, mode)
    return int(lib().nuslam_scan_set_fit(int(code)))


def last_fallbacks(device=0) -> int:
    """Scans of this thread's last moment-mode call that were re-run by the oracle-order kernel (-1: no such call)."""
    return int(lib().nuslam_scan_last_fallbacks(int(device)))


def scan_detect(ranges, min_range, max_range, max_circles=16, want_cluster_of_beam=True, device=0, stream=None):
    """clusterPoints -> classifyCluster -> circleFit -> the landmarks node's filters, for S scans of 360 float ranges.

    Returns a dict: ``cluster_of_beam`` [S,360] int16 (index of the returned cluster holding the beam, -1 none),
    ``n_clusters`` [S], ``n_circles`` [S] (``SCAN_UB`` where the reference indexes clusters[0] of an empty vector),
    ``circles`` [S,max_circles,4] = (cx, cy, R = scale.x / 2, cluster index) in detection order."""
    if _is_torch(ranges) and ranges.is_cuda:
        import torch
        if ranges.dtype != torch.float32 or not ranges.is_contiguous():
            raise NuslamError("ranges must be a contiguous float32 tensor")
        r = ranges.reshape(-1, BEAMS)
        S = r.shape[0]
        dev = r.device
        cob = torch.empty((S, BEAMS), dtype=torch.int16, device=dev) if want_cluster_of_beam else None
        ncl = torch.empty(S, dtype=torch.int32, device=dev)
        nci = torch.empty(S, dtype=torch.int32, device=dev)
        circ = torch.zeros((S, max_circles, 4), dtype=torch.float64, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        _check(lib().nuslam_scan_detect(r.data_ptr(), S, float(min_range), float(max_range), cob.data_ptr() if cob is not None else None,
                                        ncl.data_ptr(), nci.data_ptr(), circ.data_ptr(), max_circles, NUSLAM_DEVICE, dev.index, st),
               "nuslam_scan_detect")
        return dict(cluster_of_beam=cob, n_clusters=ncl, n_circles=nci, circles=circ)
    r = np.ascontiguousarray(ranges, dtype=np.float32).reshape(-1, BEAMS)
    S = r.shape[0]
    cob = np.empty((S, BEAMS), dtype=np.int16) if want_cluster_of_beam else None
    ncl = np.empty(S, dtype=np.int32)
    nci = np.empty(S, dtype=np.int32)
    circ = np.zeros((S, max_circles, 4))
    if S:
        _check(lib().nuslam_scan_detect(r.ctypes.data, S, float(min_range), float(max_range), cob.ctypes.data if cob is not None else None,
                                        ncl.ctypes.data, nci.ctypes.data, circ.ctypes.data, max_circles, NUSLAM_HOST, device, stream),
               "nuslam_scan_detect")
    return dict(cluster_of_beam=cob, n_clusters=ncl, n_circles=nci, circles=circ)


def clusterPoints(ranges, minRange, maxRange, device=0):
    """circle_fit::clusterPoints (circle_fit_library.cpp:136-206) for ONE scan: list of clusters, each a list of beam
    indices in the reference's stored order (the wrap point 359 comes last in cluster 0). Raises where the reference has
    undefined behaviour."""
    r = np.ascontiguousarray(ranges, dtype=np.float32).reshape(BEAMS)
    out = scan_detect(r[None], minRange, maxRange, device=device)
    if out["n_circles"][0] == SCAN_UB:
        raise NuslamError("clusterPoints: the reference indexes clusters[0] of an empty vector for this scan (undefined behaviour)")
    cob = out["cluster_of_beam"][0]
    clusters = []
    for k in range(int(out["n_clusters"][0])):
        # ascending beam order is the stored order: beam 359 appended by the wrap rule comes last in cluster 0
        clusters.append([int(b) for b in np.nonzero(cob == k)[0]])
    return clusters


def classify_and_fit(points_per_cluster, device=0):
    """classifyCluster + circleFit for C explicit clusters. Returns (is_circle [C] bool, marker_id [C], (cx, cy, R) [C,3])."""
    offs = np.zeros(len(points_per_cluster) + 1, dtype=np.int32)
    for k, pts in enumerate(points_per_cluster):
        offs[k + 1] = offs[k] + len(pts)
    allp = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in points_per_cluster]) if len(points_per_cluster) else np.zeros((0, 2))
    px = np.ascontiguousarray(allp[:, 0])
    py = np.ascontiguousarray(allp[:, 1])
    nC = len(points_per_cluster)
    is_c = np.zeros(nC, dtype=np.int32)
    fit = np.zeros((nC, 4))
    if nC:
        _check(lib().nuslam_classify_and_fit(px.ctypes.data, py.ctypes.data, offs.ctypes.data, nC, is_c.ctypes.data, fit.ctypes.data,
                                             NUSLAM_HOST, device, None), "nuslam_classify_and_fit")
    return is_c.astype(bool), fit[:, 0].astype(np.int32), fit[:, 1:4]


def classifyCluster(cluster, device=0) -> bool:
    """circle_fit::classifyCluster (circle_fit_library.cpp:208-250); cluster: [N,2] points."""
    return bool(classify_and_fit([cluster], device)[0][0])


def circleFit(data, device=0):
    """circle_fit::circleFit (circle_fit_library.cpp:15-134); data: [N,2] points.
    Returns (marker.id, pose.x, pose.y, R) with R = marker.scale.x / 2 (the reference stores scale.x = 2R)."""
    _, mid, f = classify_and_fit([data], device)
    return int(mid[0]), float(f[0, 0]), float(f[0, 1]), float(f[0, 2])
