"""Alias registry: build configurable objects from a string / dict / instance.

Mirrors the plugin boundary of the reference (``pydrobert/speech/alias.py:28-100``): every
configurable class (frame computers, banks, scales, windows, pre/post processors) derives from
:class:`AliasedFactory`, advertises a set of ``aliases`` and can be constructed from a JSON-like
argument with :func:`alias_factory_subclass_from_arg`.  The JSON configs accepted by the
reference are accepted unchanged.
"""

import abc
from typing import Any, Mapping, Set, Union

__all__ = ["alias_factory_subclass_from_arg", "AliasedFactory"]


def _descendants_last_registered_first(root):
    """Yield ``root``'s subclass tree so that later-registered, deeper classes come first.

    The reference resolves alias clashes by "always trying the last registered subclass that
    matches" (``alias.py:34-69``); it does so with an explicit stack that visits children after
    pushing the parent back.  The visiting order that results is: for each node, its children in
    *reverse* registration order (each fully explored), then the node itself.
    """
    for child in reversed(root.__subclasses__()):
        yield from _descendants_last_registered_first(child)
    yield root


class AliasedFactory(abc.ABC):
    """Base class of everything that can be named in a config"""

    aliases: Set[str] = set()

    @classmethod
    def from_alias(cls, alias: str, *args, **kwargs):
        """Instantiate the subclass of ``cls`` (or ``cls`` itself) registered under ``alias``

        Raises
        ------
        ValueError
            If no class in the subtree carries the alias (same message as the reference).
        """
        seen = set()
        for klass in _descendants_last_registered_first(cls):
            if klass in seen:  # diamond inheritance: visit once
                continue
            seen.add(klass)
            if alias in klass.aliases:
                return klass(*args, **kwargs)
        raise ValueError(f"Cannot find subclass with alias '{alias}'")


def alias_factory_subclass_from_arg(
    factory_class, arg: Union[AliasedFactory, str, Mapping[str, Any]]
):
    """Turn ``arg`` into an instance of ``factory_class``

    * an instance of ``factory_class`` is returned untouched;
    * a string is an alias with no arguments;
    * a mapping is copied, its ``"alias"`` key (or, failing that, ``"name"``) popped as the
      alias and the remainder passed as keyword arguments.
    """
    if isinstance(arg, factory_class):
        return arg
    if isinstance(arg, str):
        return factory_class.from_alias(arg)
    kwargs = dict(arg)
    if "alias" in kwargs:
        alias = kwargs.pop("alias")
    else:
        alias = kwargs.pop("name")
    return factory_class.from_alias(alias, **kwargs)
