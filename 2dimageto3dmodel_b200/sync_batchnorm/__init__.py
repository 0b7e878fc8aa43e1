"""Drop-in for /root/reference/code/sync_batchnorm: cross-replica batch norm for the generator's
ConditionalBatchNorm2d (models/gan.py:267-269).  The reference synchronises nn.DataParallel replicas
through Python threads/queues (sync_batchnorm/comm.py) and ReduceAddCoalesced/Broadcast
(batchnorm.py:110-131); here every rank is its own process and the statistics [sum, sum of squares] travel in
ONE NCCL all-reduce per layer (torch.distributed), forward and backward."""
from .batchnorm import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d, convert_model
from .replicate import DataParallelWithCallback, patch_replication_callback

__all__ = ['SynchronizedBatchNorm1d', 'SynchronizedBatchNorm2d', 'SynchronizedBatchNorm3d', 'convert_model',
           'DataParallelWithCallback', 'patch_replication_callback']
