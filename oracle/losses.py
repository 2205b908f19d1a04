"""NumPy restatement of tf_raft/losses/losses.py:4-43 (test infrastructure)."""
import numpy as np


def sequence_loss(y_true, y_pred, gamma=0.8, max_flow=400):
    flow_gt, valid = y_true
    flow_gt = np.asarray(flow_gt, np.float32)
    mag = np.sqrt(np.sum(flow_gt ** 2, axis=-1))
    valid = (np.asarray(valid, bool) & (mag < max_flow)).astype(np.float32)[..., None]
    n = len(y_pred)
    loss = 0.0
    for i in range(n):
        loss += gamma ** (n - i - 1) * np.mean(valid * np.abs(np.asarray(y_pred[i], np.float32) - flow_gt))
    return np.float32(loss)


def end_point_error(y_true, y_pred, max_flow=400):
    flow_gt, valid = y_true
    flow_gt = np.asarray(flow_gt, np.float32)
    mag = np.sqrt(np.sum(flow_gt ** 2, axis=-1))
    valid = np.asarray(valid, bool) & (mag < max_flow)
    epe = np.sqrt(np.sum((np.asarray(y_pred, np.float32) - flow_gt) ** 2, axis=-1))[valid]
    return {'epe': epe.mean(), 'u1': (epe < 1).mean(), 'u3': (epe < 3).mean(), 'u5': (epe < 5).mean()}
