import json

from . import new_subdir  # noqa: F401
from .new_subdir import get_new_subdir  # noqa: F401


def dumps_json(obj, *, indent=2, sort_keys=True, **_):
    return json.dumps(obj, indent=indent, sort_keys=sort_keys, default=str)


def loads_json(text):
    return json.loads(text)
