"""Attribute-compatible stand-ins for the few `qdrant_client.http.models` classes the retrievers build
(SearchParams, Prefetch, Filter, HasIdCondition, FieldCondition, MatchAny, MatchValue, MatchExcept, Range).  The reference
imports them from qdrant_client (two_stage.py:25-26, three_stage.py:76); the GPU backend only reads their
attributes, so real qdrant objects work as well."""

from __future__ import annotations


class _Kw:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):  # pragma: no cover
        return f"{type(self).__name__}({self.__dict__})"


class SearchParams(_Kw):
    pass


class Prefetch(_Kw):
    pass


class Filter(_Kw):
    pass


class HasIdCondition(_Kw):
    pass


class FieldCondition(_Kw):
    pass


class MatchAny(_Kw):
    pass


class MatchValue(_Kw):
    pass


class MatchExcept(_Kw):
    def __init__(self, **kw):
        if "except" in kw:            # qdrant's field is `except_` (alias "except")
            kw["except_"] = kw.pop("except")
        super().__init__(**kw)


class Range(_Kw):
    pass
