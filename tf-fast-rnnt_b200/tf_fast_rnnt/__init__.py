"""tf_fast_rnnt — B200 (sm_100a) drop-in for the pruned RNN-T loss hot path of
Samsung/tf-fast-rnnt.  Same public names as the reference package
(tf_fast_rnnt/python/tf_fast_rnnt/__init__.py:24-36, 42, 151); the compute is a
C-ABI library of hand-written CUDA kernels (include/fast_rnnt_b200.h)."""
from ._lib import FastRnntError, LIB_PATH  # noqa: F401  (import fails loudly without the .so)
from .rnnt_loss import cummin
from .rnnt_loss import do_rnnt_pruning
from .rnnt_loss import do_rnnt_pruning_backward
from .rnnt_loss import get_rnnt_logprobs
from .rnnt_loss import get_rnnt_logprobs_joint
from .rnnt_loss import get_rnnt_logprobs_pruned
from .rnnt_loss import get_rnnt_logprobs_smoothed
from .rnnt_loss import get_rnnt_prune_ranges
from .rnnt_loss import mutual_information_recursion
from .rnnt_loss import do_rnnt_pruning_add_joiner, pruned_add_joiner
from .rnnt_loss import pruned_loss_fwd_bwd
from .rnnt_loss import rnnt_loss
from .rnnt_loss import rnnt_loss_pruned
from .rnnt_loss import rnnt_loss_simple
from .rnnt_loss import rnnt_loss_smoothed
from .rnnt_loss import simple_loss_backward, smoothed_loss_backward
from .scheduler import make_buckets, pruned_rnnt_pipeline

__version__ = "1.2"  # the reference's version string (__init__.py:36)
