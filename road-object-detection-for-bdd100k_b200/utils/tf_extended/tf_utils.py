"""`utils/tf_extended/tf_utils.py:29-55`: the list plumbing train.py:114-124 wraps around the per-layer
outputs of refine_groundtruth (flatten a list of lists for batching, then restore the structure)."""

__all__ = ["reshape_list"]


def reshape_list(l, shape=None):
    """shape None: flatten one level of nesting.  Otherwise `shape` lists group sizes: a size of 1 keeps
    the element itself, a larger size takes that many consecutive elements as a sub-list."""
    if shape is None:
        flat = []
        for item in l:
            flat.extend(item) if isinstance(item, (list, tuple)) else flat.append(item)
        return flat
    out, pos = [], 0
    for size in shape:
        out.append(l[pos] if size == 1 else l[pos:pos + size])
        pos += size
    return out
