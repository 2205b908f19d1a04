"""Losses / metrics -- host-side mirror of tf_raft/losses/losses.py (torch elementwise ops; not on
the hot path).  `y_true = (flow_gt, valid)`, predictions NHWC."""
import torch


def _valid_mask(flow_gt, valid, max_flow):
    mag = torch.sqrt(torch.sum(flow_gt ** 2, dim=-1))
    return valid.to(torch.bool) & (mag < max_flow)


def sequence_loss(y_true, y_pred, gamma=0.8, max_flow=400):
    """Reference losses.py:4-21: gamma-weighted L1 over the prediction sequence."""
    flow_gt, valid = y_true
    flow_gt = torch.as_tensor(flow_gt, dtype=torch.float32)
    valid = _valid_mask(flow_gt, torch.as_tensor(valid, device=flow_gt.device), max_flow)
    valid = valid.to(torch.float32).unsqueeze(-1)
    n = len(y_pred)
    loss = 0.0
    for i in range(n):
        w = gamma ** (n - i - 1)
        pred = torch.as_tensor(y_pred[i], dtype=torch.float32, device=flow_gt.device)
        loss = loss + w * torch.mean(valid * torch.abs(pred - flow_gt))
    return loss


def end_point_error(y_true, y_pred, max_flow=400):
    """Reference losses.py:24-43: epe and the <1 / <3 / <5 px rates over valid pixels."""
    flow_gt, valid = y_true
    flow_gt = torch.as_tensor(flow_gt, dtype=torch.float32)
    valid = _valid_mask(flow_gt, torch.as_tensor(valid, device=flow_gt.device), max_flow)
    pred = torch.as_tensor(y_pred, dtype=torch.float32, device=flow_gt.device)
    epe = torch.sqrt(torch.sum((pred - flow_gt) ** 2, dim=-1))[valid]
    return {'epe': epe.mean(), 'u1': (epe < 1).float().mean(), 'u3': (epe < 3).float().mean(),
            'u5': (epe < 5).float().mean()}


class EndPointError:
    """Reference losses.py:46-85: the streaming metric -- per update the means of epe / <1 / <3 / <5 px over the valid
    pixels of the LAST prediction are accumulated; result() is their average over the updates."""

    def __init__(self, max_flow=400, **kwargs):
        self.max_flow = max_flow
        self.reset_states()

    def reset_states(self):
        self.epe = self.u1 = self.u3 = self.u5 = 0.0
        self.count = 0

    def update_state(self, y_true, y_pred):
        info = end_point_error(y_true, y_pred[-1], self.max_flow)
        self.epe += float(info['epe'])
        self.u1 += float(info['u1'])
        self.u3 += float(info['u3'])
        self.u5 += float(info['u5'])
        self.count += 1

    def result(self):
        n = max(self.count, 1)
        return {'epe': self.epe / n, 'u1': self.u1 / n, 'u3': self.u3 / n, 'u5': self.u5 / n}
