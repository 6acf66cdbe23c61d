def get_new_subdir(*args, **kwargs):
    raise NotImplementedError('stand-in: storage-dir allocation is outside the hot path')
