"""Identity-indexed memo tables with build rules -- the behaviour of ``stheno/lazy.py:27-168`` (universal, left and
right rules; first matching rule wins), written as plain dict lookups."""

__all__ = ["LazyVector", "LazyMatrix"]


def _index(key):
    if isinstance(key, int):
        return key
    if isinstance(key, (tuple, reversed)):
        return tuple(_index(k) for k in key)
    return id(key)


class _LazyTensor:
    def __init__(self, rank):
        self._rank = rank
        self._store = {}

    def _expand(self, key):
        return key if isinstance(key, tuple) else (key,) * self._rank

    def __setitem__(self, key, value):
        self._store[_index(self._expand(key))] = value

    def __getitem__(self, key):
        i = _index(self._expand(key))
        try:
            return self._store[i]
        except KeyError:
            pass
        value = self._build(i)
        self._store[i] = value
        return value


class LazyVector(_LazyTensor):
    def __init__(self):
        super().__init__(1)
        self._rules = []

    def add_rule(self, indices, builder):
        self._rules.append((frozenset(indices), builder))

    def _build(self, i):
        (i,) = i
        for indices, builder in self._rules:
            if i in indices:
                return builder(i)
        raise RuntimeError(f'Could not build value for index "{i}".')


class LazyMatrix(_LazyTensor):
    def __init__(self):
        super().__init__(2)
        self._left_rules, self._right_rules, self._rules = [], [], []

    def add_rule(self, indices, builder):
        self._rules.append((frozenset(indices), builder))

    def add_left_rule(self, i_left, indices, builder):
        self._left_rules.append((i_left, frozenset(indices), builder))

    def add_right_rule(self, i_right, indices, builder):
        self._right_rules.append((i_right, frozenset(indices), builder))

    def _build(self, i):
        i_left, i_right = i
        for indices, builder in self._rules:
            if i_left in indices and i_right in indices:
                return builder(i_left, i_right)
        for i_left_rule, indices, builder in self._left_rules:
            if i_left == i_left_rule and i_right in indices:
                return builder(i_right)
        for i_right_rule, indices, builder in self._right_rules:
            if i_left in indices and i_right == i_right_rule:
                return builder(i_left)
        raise RuntimeError(f"Could not build value for index {i}.")
