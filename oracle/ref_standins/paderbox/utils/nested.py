"""Tree utilities over dict / list / tuple (/ dataclass) containers."""
import copy
import dataclasses


def flatten(d, sep='.', *, flat_type=dict):
    flat = {}

    def walk(path, node):
        if isinstance(node, flat_type) and len(node):
            for key, child in node.items():
                walk(path + (key,), child)
        else:
            flat[path if sep is None else sep.join(str(p) for p in path)] = node

    walk((), d)
    return flat


def deflatten(d, sep='.', maxdepth=-1):
    tree = {}
    for key, value in d.items():
        path = list(key) if sep is None else key.split(sep, maxdepth)
        node = tree
        for part in path[:-1]:
            node = node.setdefault(part, {})
        node[path[-1]] = value
    return tree


def nested_op(func, arg1, *args, broadcast=False, handle_dataclass=False,
              keep_type=True, mapping_type=dict, sequence_type=(tuple, list)):
    kw = dict(handle_dataclass=handle_dataclass, mapping_type=mapping_type,
              sequence_type=sequence_type)
    if isinstance(arg1, mapping_type):
        out = {k: nested_op(func, v, *[a[k] for a in args], **kw) for k, v in arg1.items()}
        return arg1.__class__(out) if keep_type else out
    if isinstance(arg1, sequence_type):
        out = [nested_op(func, *items, **kw) for items in zip(arg1, *args)]
        return arg1.__class__(out) if keep_type else out
    if handle_dataclass and dataclasses.is_dataclass(arg1) and not isinstance(arg1, type):
        fields = {f.name: nested_op(func, getattr(arg1, f.name),
                                    *[getattr(a, f.name) for a in args], **kw)
                  for f in dataclasses.fields(arg1)}
        return arg1.__class__(**fields)
    return func(arg1, *args)


def nested_merge(base, *updates, **_):
    merged = copy.copy(base)
    for update in updates:
        for key, value in update.items():
            if isinstance(value, dict) and isinstance(merged.get(key), dict):
                merged[key] = nested_merge(merged[key], value)
            else:
                merged[key] = value
    return merged
