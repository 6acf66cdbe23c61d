from . import database  # noqa: F401


class FilterException(Exception):
    pass


def from_list(examples):
    return list(examples)
