"""Drop-in for /root/reference/code/utils/batch_repetition.py:6-19: every batch element repeated n times in place
([a, b] -> [a, a, a, b, b, b]) — how masks / point clouds are matched to the K pose candidates and V views."""
import torch


def repeat_tensor_for_each_element_in_batch(torch_tensor, n):
    return torch.repeat_interleave(torch_tensor, int(n), dim=0)
