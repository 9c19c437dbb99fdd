"""[B,V,...] <-> [B*V,...] reshapes.  API mirror of reference ``latentfusion/three/batchview.py``."""
import torch


def bv2b(x):
    return x.reshape(-1, *x.shape[2:])


def b2bv(x, num_view=-1, batch_size=-1):
    if num_view == -1 and batch_size == -1:
        raise ValueError('One of num_view or batch_size must be non-negative.')
    return x.reshape(batch_size, num_view, *x.shape[1:])


def bvmm(a, b):
    if a.shape[:2] != b.shape[:2]:
        raise ValueError("batch and view dimensions must match")
    return b2bv(torch.bmm(bv2b(a), bv2b(b)), num_view=a.shape[1])


def vcat(tensors, batch_size):
    return bv2b(torch.cat([b2bv(t, batch_size=batch_size) for t in tensors], dim=1))


def vsplit(tensor, sections):
    t = b2bv(tensor, num_view=sum(sections))
    return tuple(bv2b(s) for s in torch.split(t, sections, dim=1))
