from typing import (IO, Any, Callable, Generator, Iterable, Iterator, Literal, NoReturn, Protocol, Sequence, TypedDict,
                    Union, overload, runtime_checkable)

__all__ = ["IO", "Any", "Callable", "Generator", "Iterable", "Iterator", "Literal", "NoReturn", "Protocol", "Sequence",
           "TypedDict", "Union", "overload", "runtime_checkable"]
