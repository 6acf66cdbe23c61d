import difflib


class DispatchError(KeyError):
    pass


class Dispatcher(dict):
    """dict that reports close matches on a missing key."""

    def __getitem__(self, key):
        if key in self:
            return dict.__getitem__(self, key)
        close = difflib.get_close_matches(str(key), [str(k) for k in self])
        raise DispatchError(f'Invalid option {key!r}. Close matches: {close}; all: {list(self)}')
