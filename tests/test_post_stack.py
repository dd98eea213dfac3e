"""post.Stack (reference post.py:494-563) against outputs of the real reference (extra.npz); the
processor is a re-indexing done on the host, so this runs without a GPU."""
import numpy as np
import pytest

import cases


@pytest.mark.parametrize("name,kwargs,axis", cases.STACK_CASES)
def test_stack_matches_reference(speech, golden, name, kwargs, axis):
    data = golden("extra")
    got = speech.post.Stack(**kwargs).apply(data["stack/feats"], axis=axis)
    want = data["stack/" + name]
    assert got.shape == want.shape and np.array_equal(got, want)


def test_stack_errors_and_alias(speech):
    with pytest.raises(ValueError):
        speech.post.Stack(0)
    stack = speech.alias_factory_subclass_from_arg(speech.post.PostProcessor, {"name": "stack", "num_vectors": 2})
    with pytest.raises(RuntimeError):
        stack.apply(np.zeros((4, 3)), axis=0)  # feature axis == time axis
    assert stack.apply(np.arange(12.0).reshape(4, 3)).shape == (2, 6)
