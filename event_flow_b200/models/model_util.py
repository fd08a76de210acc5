"""State-copy helpers with the semantics of models/model_util.py:82-102 (model.states returns clones)."""
import copy


def recursive_clone(tensor):
    if hasattr(tensor, "clone"):
        return tensor.clone()
    try:
        return type(tensor)(recursive_clone(t) for t in tensor)
    except TypeError:
        print("{} is not iterable and has no clone() method.".format(tensor))


def copy_states(states):
    if states[0] is None:
        return copy.deepcopy(states)
    return recursive_clone(states)
