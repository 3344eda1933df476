"""Drop-in `model.loss` surface: MaxMarginRankingLoss (reference model/loss.py:29-65) on the
sm_100a max-margin kernel (forward and gradient in one streaming pass over the N x N matrix)."""
import torch
import torch.nn as nn

from .. import engine


class MaxMarginFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, margin, fix_norm):
    x = x.contiguous()
    loss, dx = engine.max_margin(x, margin, fix_norm, want_grad=True)
    ctx.save_for_backward(dx)
    return loss

  @staticmethod
  def backward(ctx, g):
    dx, = ctx.saved_tensors
    return dx * g, None, None


class MaxMarginRankingLoss(nn.Module):
  """Implementation of the Max-margin ranking loss (same ctor / call as the reference)."""

  def __init__(self, margin=1, fix_norm=True):
    super().__init__()
    self.fix_norm = fix_norm
    self.margin = margin

  def forward(self, x):
    if x.dim() != 2 or x.size(0) != x.size(1):
      raise ValueError("MaxMarginRankingLoss expects a square similarity matrix")
    in_dev = x.device
    if not x.is_cuda:                       # no CPU arithmetic: compute on the GPU, return on input device
      x = x.to(torch.device("cuda", torch.cuda.current_device()))
    loss = MaxMarginFn.apply(x.to(torch.float32), float(self.margin), bool(self.fix_norm))
    return loss.to(in_dev)
