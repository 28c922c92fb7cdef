"""Normal-estimation metrics of the reference's utils/evaluate.py (:105-200), SURVEY.md 8(f) rank 3:
unoriented / oriented RMS angle and the proportion of good points (PGP) at 5 and 10 degrees.
Pure numpy, no device -- the same definitions the parity report of the MoE normals uses."""
import numpy as np


def normalize_rows(n):
    """utils/evaluate.py:128-131: divide every normal by its norm (zero rows stay zero)."""
    n = np.asarray(n, dtype=np.float64)
    norm = np.sqrt((n ** 2).sum(axis=1, keepdims=True))
    return np.divide(n, norm, out=np.zeros_like(n), where=norm > 0)


def angle_errors_deg(normals_pred, normals_gt, oriented=False):
    """Per-point angle in degrees between predicted and ground-truth normals (:139-147).
    Unoriented: arccos(|n . n_gt|); oriented: arccos(n . n_gt)."""
    a = normalize_rows(normals_pred)
    b = normalize_rows(normals_gt)
    dot = np.clip((a * b).sum(axis=1), -1.0, 1.0)
    if not oriented:
        dot = np.abs(dot)
    return np.rad2deg(np.arccos(dot))


def rms_angle(normals_pred, normals_gt, oriented=False):
    """sqrt(mean(angle^2)) in degrees (:148-150)."""
    ang = angle_errors_deg(normals_pred, normals_gt, oriented)
    return float(np.sqrt(np.mean(ang ** 2)))


def pgp(normals_pred, normals_gt, threshold_deg):
    """Proportion of good points: fraction of unoriented angle errors below the threshold (:151-154)."""
    ang = angle_errors_deg(normals_pred, normals_gt, oriented=False)
    return float(np.mean(ang < threshold_deg))


def evaluate_shape(normals_pred, normals_gt, pidx=None):
    """The per-shape record of utils/evaluate.py: rms (unoriented), rms_o (oriented), pgp5, pgp10.
    `pidx`: optional sparse subset of point indices (the .pidx files, :118-124)."""
    pred = np.asarray(normals_pred)
    gt = np.asarray(normals_gt)
    if pidx is not None:
        gt = gt[np.asarray(pidx, dtype=np.int64)]
        if len(pred) != len(gt):
            pred = pred[np.asarray(pidx, dtype=np.int64)]
    return {"rms": rms_angle(pred, gt), "rms_o": rms_angle(pred, gt, oriented=True),
            "pgp5": pgp(pred, gt, 5.0), "pgp10": pgp(pred, gt, 10.0), "n": int(len(gt))}
