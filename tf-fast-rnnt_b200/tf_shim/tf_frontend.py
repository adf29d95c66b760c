"""TensorFlow front-end over the shim ops (tf_fast_rnnt_b200_ops.cc).

Drop this in place of the reference's ``tf_fast_rnnt/__init__.py`` + ``rnnt_loss.py``
on a machine that has TensorFlow: same public names and signatures, the graph
math replaced by the fused ops.  It cannot be imported in this repository's
container (TensorFlow is absent); the TensorFlow-free twin with identical
semantics, which the tests exercise, is ``tf_fast_rnnt/rnnt_loss.py``.
"""
import glob
import os

import tensorflow as tf
from tensorflow.python.framework import ops

_here = os.path.dirname(os.path.abspath(__file__))
_so = glob.glob(os.path.join(_here, "_tf_fast_rnnt*.so"))       # reference: imp.find_module (__init__.py:38-40)
if not _so:
    raise ImportError("_tf_fast_rnnt*.so not found next to tf_frontend.py; build tf_fast_rnnt_b200_ops.cc")
_ops = tf.load_op_library(_so[0])
_TYPES = {"regular": 0, "modified": 1, "constrained": 2}
__version__ = "1.2"


def mutual_information_recursion(px, py, boundary, calc_gradients=False):
    ans, gx, gy = _ops.fast_rnnt_loss(px, py, boundary, calc_gradients)
    return (ans, (gx, gy)) if calc_gradients else ans


def cummin(x):
    return _ops.cummin(x)


@ops.RegisterGradient("FastRNNTLoss")
def _rnnt_loss_grad(op, *grads):           # reference: __init__.py:154-162
    g = tf.reshape(grads[0], (-1, 1, 1))
    return [g * op.outputs[1], g * op.outputs[2], None, None]


@ops.RegisterGradient("FastRnntDoPruning")
def _do_pruning_grad(op, g_am, g_lm):
    S = op.inputs[1].shape[1] - 1
    am_g, lm_g = _ops.fast_rnnt_do_pruning_grad(g_am, g_lm, op.inputs[2], S=S)
    return [am_g, lm_g, None]


@ops.RegisterGradient("FastRnntDoPruningAddJoiner")
def _do_pruning_add_joiner_grad(op, g_am, g_lm, g_logits):
    S = op.inputs[1].shape[1] - 1
    am_g, lm_g = _ops.fast_rnnt_do_pruning_grad(g_am + g_logits, g_lm + g_logits, op.inputs[2], S=S)
    return [am_g, lm_g, None]


@ops.RegisterGradient("FastRnntSimpleLoss")
def _simple_loss_grad(op, g_scores, _g_px_grad, _g_py_grad):
    """A9: d scores / d (lm, am) from the occupation counts the forward op produced (it must have
    run with calc_gradients=True).  The forward op takes bf16 / fp16 lm, am as they are; the gradient op is
    float32: low-precision inputs are widened here and the gradients returned in the inputs' type."""
    in_type = op.inputs[0].dtype
    lm_g, am_g = _ops.fast_rnnt_simple_loss_grad(
        tf.cast(op.inputs[0], tf.float32), tf.cast(op.inputs[1], tf.float32), op.inputs[2], op.inputs[3],
        op.outputs[1], op.outputs[2], g_scores,
        termination_symbol=op.get_attr("termination_symbol"), rnnt_type=op.get_attr("rnnt_type"),
        smoothed=op.get_attr("smoothed"), lm_only_scale=op.get_attr("lm_only_scale"),
        am_only_scale=op.get_attr("am_only_scale"))
    return [tf.cast(lm_g, in_type), tf.cast(am_g, in_type), None, None]


@ops.RegisterGradient("FastRnntSimpleLogprobs")
def _simple_logprobs_grad(op, g_px, g_py):
    """Backward of get_rnnt_logprobs{,_smoothed}: the contractions of the loss gradient, fed with the
    cotangents of px / py (the reference gets this from TF autodiff over rnnt_loss.py:175-221)."""
    ones = tf.ones([tf.shape(op.inputs[1])[0]], tf.float32)
    lm_g, am_g = _ops.fast_rnnt_simple_loss_grad(
        op.inputs[0], op.inputs[1], op.inputs[2], op.inputs[3], g_px, g_py, ones,
        termination_symbol=op.get_attr("termination_symbol"), rnnt_type=op.get_attr("rnnt_type"),
        smoothed=op.get_attr("smoothed"), lm_only_scale=op.get_attr("lm_only_scale"),
        am_only_scale=op.get_attr("am_only_scale"))
    return [lm_g, am_g, None, None]


@ops.RegisterGradient("FastRnntPrunedLogprobs")
def _pruned_logprobs_grad(op, g_px, g_py):
    """Backward of get_rnnt_logprobs_pruned (the reference composes it with mutual_information_recursion under
    autodiff, rnnt_loss.py:1088-1119)."""
    g = _ops.fast_rnnt_pruned_logprobs_grad(op.inputs[0], op.inputs[1], op.inputs[2], op.inputs[3], g_px, g_py,
                                            termination_symbol=op.get_attr("termination_symbol"),
                                            rnnt_type=op.get_attr("rnnt_type"))
    return [g, None, None, None]


def _reduce(scores, reduction):
    if reduction == "none":
        return -scores
    if reduction == "mean":
        return -tf.reduce_mean(scores)
    if reduction == "sum":
        return -tf.reduce_sum(scores)
    raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")


def _boundary(boundary, B, S, T):
    if boundary is None:
        return tf.tile(tf.constant([[0, 0, S, T]], tf.int32), [B, 1])
    return tf.cast(boundary, tf.int32)


def _simple(lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty, reduction,
            calc_gradients, smoothed, lms, ams):
    B, T = am.shape[0], am.shape[1]
    S = lm.shape[1] - 1
    scores, gx, gy = _ops.fast_rnnt_simple_loss(
        lm, am, tf.cast(symbols, tf.int32), _boundary(boundary, B, S, T), termination_symbol=termination_symbol,
        rnnt_type=_TYPES[rnnt_type], smoothed=smoothed, lm_only_scale=lms, am_only_scale=ams,
        delay_penalty=max(delay_penalty, 0.0), calc_gradients=True)      # the occupation counts feed the gradient op
    loss = _reduce(scores, reduction)
    return (loss, (gx, gy)) if calc_gradients else loss


def rnnt_loss_simple(lm, am, symbols, termination_symbol, boundary=None, rnnt_type="regular",
                     delay_penalty=0.0, reduction="mean", calc_gradients=False):
    return _simple(lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty, reduction,
                   calc_gradients, False, 0.0, 0.0)


def rnnt_loss_smoothed(lm, am, symbols, termination_symbol, lm_only_scale=0.1, am_only_scale=0.1,
                       boundary=None, rnnt_type="regular", delay_penalty=0.0, reduction="mean",
                       calc_gradients=False):
    return _simple(lm, am, symbols, termination_symbol, boundary, rnnt_type, delay_penalty, reduction,
                   calc_gradients, True, lm_only_scale, am_only_scale)


def get_rnnt_prune_ranges(px_grad, py_grad, boundary, s_range):
    return _ops.fast_rnnt_prune_ranges(px_grad, py_grad, tf.cast(boundary, tf.int32), s_range=s_range)


def do_rnnt_pruning(am, lm, ranges):
    return _ops.fast_rnnt_do_pruning(am, lm, ranges)


def do_rnnt_pruning_add_joiner(am, lm, ranges):
    """(extension) -> (am_pruned, lm_pruned, am_pruned + lm_pruned) in one pass."""
    return _ops.fast_rnnt_do_pruning_add_joiner(am, lm, ranges)


def rnnt_loss_pruned(logits, symbols, ranges, termination_symbol, boundary=None, rnnt_type="regular",
                     delay_penalty=0.0, reduction="mean", calc_gradients=False):
    B, T = logits.shape[0], logits.shape[1]
    S = symbols.shape[1]
    bd = _boundary(boundary, B, S, T)
    sym = tf.cast(symbols, tf.int32)

    @tf.custom_gradient
    def _scores(lg):
        s, _ = _ops.fast_rnnt_pruned_loss(lg, sym, ranges, bd, tf.ones([B], tf.float32),
                                          termination_symbol=termination_symbol, rnnt_type=_TYPES[rnnt_type],
                                          delay_penalty=max(delay_penalty, 0.0), with_logits_grad=False)

        def grad(upstream):
            _, g = _ops.fast_rnnt_pruned_loss(lg, sym, ranges, bd, upstream,
                                              termination_symbol=termination_symbol, rnnt_type=_TYPES[rnnt_type],
                                              delay_penalty=max(delay_penalty, 0.0), with_logits_grad=True)
            return g
        return s, grad

    return _reduce(_scores(logits), reduction)


def rnnt_loss(logits, symbols, termination_symbol, boundary=None, rnnt_type="regular", delay_penalty=0.0,
              reduction="mean"):
    """Reference: rnnt_loss.py:446-551 (full joiner output [B,T,S+1,C])."""
    B, T = logits.shape[0], logits.shape[1]
    S = symbols.shape[1]
    bd = _boundary(boundary, B, S, T)
    sym = tf.cast(symbols, tf.int32)
    kw = dict(termination_symbol=termination_symbol, rnnt_type=_TYPES[rnnt_type], delay_penalty=max(delay_penalty, 0.0))

    @tf.custom_gradient
    def _scores(lg):
        s, _ = _ops.fast_rnnt_joint_loss(lg, sym, bd, tf.ones([B], tf.float32), with_logits_grad=False, **kw)

        def grad(upstream):
            _, g = _ops.fast_rnnt_joint_loss(lg, sym, bd, upstream, with_logits_grad=True, **kw)
            return g
        return s, grad

    return _reduce(_scores(logits), reduction)


def get_rnnt_logprobs(lm, am, symbols, termination_symbol, rnnt_type="regular", boundary=None):
    """Reference: rnnt_loss.py:63-223 (argument order as there)."""
    B, T = am.shape[0], am.shape[1]
    S = lm.shape[1] - 1
    return _ops.fast_rnnt_simple_logprobs(lm, am, tf.cast(symbols, tf.int32), _boundary(boundary, B, S, T),
                                          termination_symbol=termination_symbol, rnnt_type=_TYPES[rnnt_type])


def get_rnnt_logprobs_smoothed(lm, am, symbols, termination_symbol, lm_only_scale=0.1, am_only_scale=0.1,
                               boundary=None, rnnt_type="regular"):
    """Reference: rnnt_loss.py:1132-1367."""
    B, T = am.shape[0], am.shape[1]
    S = lm.shape[1] - 1
    return _ops.fast_rnnt_simple_logprobs(lm, am, tf.cast(symbols, tf.int32), _boundary(boundary, B, S, T),
                                          termination_symbol=termination_symbol, rnnt_type=_TYPES[rnnt_type],
                                          smoothed=True, lm_only_scale=lm_only_scale, am_only_scale=am_only_scale)


def get_rnnt_logprobs_pruned(logits, symbols, ranges, termination_symbol, boundary, rnnt_type="regular"):
    """Reference: rnnt_loss.py:853-1020."""
    return _ops.fast_rnnt_pruned_logprobs(logits, tf.cast(symbols, tf.int32), ranges, tf.cast(boundary, tf.int32),
                                          termination_symbol=termination_symbol, rnnt_type=_TYPES[rnnt_type])


def get_rnnt_logprobs_joint(logits, symbols, termination_symbol, boundary=None, rnnt_type="regular"):
    """Reference: rnnt_loss.py:340-444: the pruned log-probs with the identity band ranges[b,t,i] = i."""
    B, T, S1 = logits.shape[0], logits.shape[1], logits.shape[2]
    ranges = tf.tile(tf.reshape(tf.range(S1, dtype=tf.int32), [1, 1, S1]), [B, T, 1])
    return get_rnnt_logprobs_pruned(logits, symbols, ranges, termination_symbol, _boundary(boundary, B, S1 - 1, T),
                                    rnnt_type)


def pruned_rnnt_pipeline(lm, am, symbols, termination_symbol, boundary, s_range, joiner=None, rnnt_type="regular",
                         delay_penalty=0.0, reduction="sum", max_buckets=8, min_bucket=4, lm_only_scale=0.0,
                         am_only_scale=0.0):
    """(extension) The full pruned RNN-T step on a ragged batch, per length bucket: the batch is
    cut into at most ``max_buckets`` buckets by frame count (``sharding.plan_buckets``: the plan is made on the
    host, so ``boundary`` must be a host-readable value - eager tensor or NumPy array), every bucket runs
    rnnt_loss_simple/_smoothed -> get_rnnt_prune_ranges -> do_rnnt_pruning -> joiner -> rnnt_loss_pruned
    trimmed to its own (S_max, T_max), and the per-utterance losses go back in batch order.  Mirrors
    tf_fast_rnnt/scheduler.py of the TensorFlow-free twin.  Returns (simple_loss, pruned_loss)."""
    try:
        from .sharding import plan_buckets
    except ImportError:                      # tf_frontend.py deployed as a plain module next to sharding.py
        from sharding import plan_buckets
    import numpy as np
    bd_host = np.asarray(boundary)
    B, C = am.shape[0], am.shape[2]
    bd = tf.cast(boundary, tf.int32)
    sym = tf.cast(symbols, tf.int32)
    simple_parts, pruned_parts, order = [], [], []
    for bk in plan_buckets(bd_host, s_range, int(C), max_buckets=max_buckets, min_bucket=min_bucket):
        idx = tf.constant(bk["idx"], tf.int32)
        lm_k = tf.gather(lm, idx)[:, :bk["S_max"] + 1]
        am_k = tf.gather(am, idx)[:, :bk["T_max"]]
        sym_k = tf.gather(sym, idx)[:, :bk["S_max"]]
        bd_k = tf.gather(bd, idx)
        if lm_only_scale > 0.0 or am_only_scale > 0.0:
            sl, (gx, gy) = rnnt_loss_smoothed(lm_k, am_k, sym_k, termination_symbol, lm_only_scale, am_only_scale, bd_k,
                                              rnnt_type, delay_penalty, "none", True)
        else:
            sl, (gx, gy) = rnnt_loss_simple(lm_k, am_k, sym_k, termination_symbol, bd_k, rnnt_type, delay_penalty,
                                            "none", True)
        ranges = get_rnnt_prune_ranges(gx, gy, bd_k, s_range)
        if joiner is None:
            _, _, logits = do_rnnt_pruning_add_joiner(am_k, lm_k, ranges)
        else:
            logits = joiner(*do_rnnt_pruning(am_k, lm_k, ranges))
        pl = rnnt_loss_pruned(logits, sym_k, ranges, termination_symbol, bd_k, rnnt_type, delay_penalty, "none")
        simple_parts.append(sl); pruned_parts.append(pl); order.append(bk["idx"])
    inv = tf.constant(np.argsort(np.concatenate(order)), tf.int32)            # back to batch order
    simple = -tf.gather(tf.concat(simple_parts, 0), inv)
    pruned = -tf.gather(tf.concat(pruned_parts, 0), inv)
    return _reduce(simple, reduction), _reduce(pruned, reduction)
