"""Container-only import shim so tests/golden/make_golden.py can import the reference from
/root/reference when the real `toolz` is absent.  Functional helpers only - no arithmetic."""
import functools
import itertools


def memoize(func=None, cache=None, key=None):
    if func is None:
        return functools.partial(memoize, cache=cache, key=key)
    store = {} if cache is None else cache

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        if key is not None:
            k = key(args, kwargs)
        elif kwargs:
            k = (args, frozenset(kwargs.items()))
        else:
            k = args
        try:
            return store[k]
        except KeyError:
            store[k] = result = func(*args, **kwargs)
            return result
        except TypeError:
            return func(*args, **kwargs)
    return wrapper


def unique(seq, key=None):
    seen = set()
    for item in seq:
        val = item if key is None else key(item)
        if val not in seen:
            seen.add(val)
            yield item


def concat(seqs):
    return itertools.chain.from_iterable(seqs)


def pluck(ind, seqs, default=None):
    if isinstance(ind, list):
        return (tuple(s[i] for i in ind) for s in seqs)
    return (s[ind] for s in seqs)


def get(ind, seq, default=None):
    if isinstance(ind, list):
        return tuple(seq[i] for i in ind)
    return seq[ind]


def identity(x):
    return x


def reduce(func, seq, *initial):
    return functools.reduce(func, seq, *initial)


def flip(func=None, a=None, b=None):
    if func is None:
        return flip
    if a is None:
        return lambda a_, b_: func(b_, a_)
    if b is None:
        return lambda b_: func(b_, a)
    return func(b, a)
