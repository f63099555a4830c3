"""CPU, build container only: oracle/git_oracle.py against the reference's own modules imported
from /root/reference (skipped where the reference tree is absent, e.g. on the GPU box)."""
import pytest
import torch

import ref_shim
import git_oracle
from generativeimage2text_b200.synthetic import state_spec, synthetic_state_dict, synthetic_images

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason='no /root/reference')


def test_state_dict_layout_matches_reference():
    for param in ({}, {'num_image_with_embedding': 6}):
        ref = ref_shim.load_reference_model(param, 'greedy', 40)
        rsd = ref.state_dict()
        spec = state_spec(param)
        assert [k for k, _, _ in spec] == list(rsd.keys())
        for k, shp, _ in spec:
            assert tuple(rsd[k].shape) == shp, k
        assert rsd['textual.output.weight'].data_ptr() == rsd['textual.embedding.words.weight'].data_ptr()


@pytest.mark.parametrize('search', ['greedy', 'beam'])
def test_oracle_equals_reference_fresh_seed(search):
    """A seed/image set that is NOT in tests/golden: both implementations run here."""
    sd = synthetic_state_dict({}, seed=7, variant='init')
    img = synthetic_images(1, 0, seed=99)
    ref = ref_shim.load_reference_model({}, search, 10, state_dict=sd)
    with torch.no_grad():
        r = ref({'image': img})
    for cached in (True, False):
        o = git_oracle.generate(sd, {}, {'image': img}, search, 10, cached=cached)
        assert torch.equal(r['predictions'], o['predictions'])
        assert torch.allclose(r['logprobs'], o['logprobs'], atol=1e-3)
