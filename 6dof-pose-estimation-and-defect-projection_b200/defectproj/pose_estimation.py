"""Point-to-plane ICP refinement: the producer of the pose the back-projection consumes.

Mirrors the ICP call sites of /root/reference/src/pose_estimation.py -- ``refine_registration`` (:505-522),
the restart loop of ``improve_result`` (:547-620) and the one-iteration probes of ``predict_z_axis_adjustment``
(:654-660) -- with Open3D's ``registration_icp`` replaced by ``dp_icp_point_to_plane`` (csrc/icp.cu): exact
nearest neighbours and the 6x6 normal equations on the GPU, the solve on the host.  Only the ICP inner loop is on
this row, plus the normals its target needs (``estimate_normals`` :301-306 -> ``dp_estimate_normals``); RANSAC/FPFH
global registration and FoundationPose stay out of scope.

Point clouds are duck-typed: anything with ``.points`` (and ``.normals`` for the target), or plain arrays.
"""
from __future__ import annotations

import copy
import logging

import numpy as np

from .defect_projection import KDTreeSearchParamHybrid, get_context

__all__ = ["ICPConvergenceCriteria", "TransformationEstimationPointToPlane", "RegistrationResult", "registration_icp",
           "refine_registration", "improve_result", "get_rotation_matrix_from_xyz", "estimate_normals"]


class ICPConvergenceCriteria:
    """o3d.pipelines.registration.ICPConvergenceCriteria (Open3D defaults)."""

    def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
        self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration


class TransformationEstimationPointToPlane:
    """Marker for the one estimation method the reference uses on this path."""


class RegistrationResult:
    def __init__(self, transformation=None, fitness=0.0, inlier_rmse=0.0, correspondence_set=None, iterations=0):
        self.transformation = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
        self.fitness, self.inlier_rmse = fitness, inlier_rmse
        self.correspondence_set = np.zeros((0, 2), np.int32) if correspondence_set is None else correspondence_set
        self.iterations = iterations

    def __repr__(self):
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, "
                f"and correspondence_set size of {len(self.correspondence_set)}")


def estimate_normals(pcd, params=None):
    """Normals of a cloud with the reference's parameters (:301-306: radius 2, at most 5 neighbours; ``params`` is
    accepted and ignored like there)."""
    pcd.estimate_normals(search_param=KDTreeSearchParamHybrid(radius=2, max_nn=5))
    return pcd


def _points(pcd):
    return np.asarray(getattr(pcd, "points", pcd), dtype=np.float64).reshape(-1, 3)


def registration_icp(source, target, max_correspondence_distance, init=None, estimation_method=None, criteria=None):
    """Same positional signature as o3d.pipelines.registration.registration_icp."""
    if estimation_method is not None and not isinstance(estimation_method, TransformationEstimationPointToPlane):
        raise ValueError("only TransformationEstimationPointToPlane is implemented (the method the reference uses)")
    criteria = criteria or ICPConvergenceCriteria()
    normals = np.asarray(getattr(target, "normals", np.zeros((0, 3))), dtype=np.float64).reshape(-1, 3)
    tp = _points(target)
    if len(normals) != len(tp):
        raise RuntimeError("TransformationEstimationPointToPlane requires target normals")
    r = get_context().icp_point_to_plane(_points(source), tp, normals, max_correspondence_distance,
                                         np.eye(4) if init is None else init, criteria.max_iteration,
                                         criteria.relative_fitness, criteria.relative_rmse, want_correspondence=True)
    c = r["correspondence"]
    idx = np.nonzero(c >= 0)[0]
    corr = np.stack([idx, c[idx]], axis=1).astype(np.int32)
    return RegistrationResult(r["transformation"], r["fitness"], r["inlier_rmse"], corr, r["iterations"])


def refine_registration(source, target, transformation, param):
    """src/pose_estimation.py:505-522."""
    params = param["refine_registration"]
    return registration_icp(source, target, params["distance_threshold"], transformation,
                            TransformationEstimationPointToPlane())


def get_rotation_matrix_from_xyz(angles):
    """o3d.geometry.get_rotation_matrix_from_xyz: R = Rx(a) Ry(b) Rz(c)."""
    a, b, c = (float(v) for v in angles)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def improve_result(source_processed, original_target_processed, current_result, parameter, rng=None, max_iterations=50):
    """The restart loop of src/pose_estimation.py:547-620: up to 50 ICP runs from jittered starts with a jittered
    distance threshold, keeping the best (fitness, then rmse).  ``rng`` (np.random.Generator) replaces the global
    np.random the reference draws from; draws are made in the reference's order."""
    rng = rng or np.random.default_rng()
    parameters = copy.deepcopy(parameter)
    if not hasattr(current_result, "fitness") or current_result.fitness is None:
        current_result = RegistrationResult(current_result, 0.8, 3.0)
    best_fitness, best_rmse = current_result.fitness, current_result.inlier_rmse
    best_transformation = np.linalg.inv(current_result.transformation)
    iteration, x = 0, 0.1
    while iteration < max_iterations and (best_fitness < parameters["run_icp"]["fitness_threshold"] or
                                          best_rmse > parameters["run_icp"]["rmse_threshold"]):
        current_param = parameters.copy()                      # shallow, like the reference: the jitter compounds
        current_param["refine_registration"]["distance_threshold"] *= rng.uniform(0.8, 1.2)
        noise_transform = np.eye(4)
        noise_transform[:3, :3] = get_rotation_matrix_from_xyz([rng.uniform(-0.01, 0.01) for _ in range(3)])
        noise_transform[:3, 3] = rng.uniform(-x, x, 3)
        current_transform = noise_transform @ best_transformation
        refined = refine_registration(source_processed, original_target_processed, current_transform, current_param)
        if refined.fitness > 0 and refined.inlier_rmse > 0:
            if refined.fitness > best_fitness or (refined.fitness == best_fitness and refined.inlier_rmse < best_rmse):
                best_fitness, best_rmse, best_transformation = refined.fitness, refined.inlier_rmse, refined.transformation
                logging.info(f":: Improved result: Fitness = {best_fitness:.4f}, RMSE = {best_rmse:.4f}")
        else:
            x += .25
        iteration += 1
    return RegistrationResult(best_transformation, best_fitness, best_rmse, iterations=iteration)
