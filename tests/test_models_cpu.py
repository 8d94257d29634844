"""CPU: the drop-in nn.Modules construct without a GPU, expose exactly the reference's
state_dict (names, shapes, dtypes, order) -- checked against the manifest dumped from the
reference classes -- load checkpoints strictly, draw the same initial weights for the same seed,
and refuse (loudly) to run on the CPU."""
import json
import os

import pytest
import torch

from evfly_b200 import _lib, learner_models as LM, vitfly_models as VM
from evfly_b200.ConvLSTM_pytorch.convlstm import ConvLSTM
from oracle.synth_ckpt import shapes_of, synth_state_dict

quiet = lambda *a, **k: None
ENC = dict(num_layers=2, kernel_sizes=[5, 3], kernel_strides=[2, 2], out_channels=[8, 32], activations=["relu", "relu"],
           pool_type="max", invert_pool_inputs=True, pool_kernels=[2, 2], pool_strides=[2, 2], conv_function="conv2d")
FC = dict(num_layers=4, layer_sizes=[1024, 128, 16, 1], activations=["leaky_relu"] * 3 + ["tanh"], dropout_p=0.1)


def build(name):
    if name in ("LSTMNetVIT", "ViT", "ConvNet", "LSTMNet", "UNetConvLSTMNet"):
        return getattr(VM, name)()
    if name == "OrigUNet_deployed":
        return LM.OrigUNet(num_in_channels=2, num_out_channels=1, num_recurrent=[1, 0], input_shape=[1, 1, 260, 346], logger=quiet,
                           velpred=0, enc_params=ENC, fc_params=FC, form_BEV=2, evs_min_cutoff=1e-3, skip_type="interp")
    if name == "OrigUNet_default":
        return LM.OrigUNet(num_recurrent=[0, 0], logger=quiet)
    if name == "OrigUNet_velpred11":
        return LM.OrigUNet(num_in_channels=2, num_recurrent=[0, 0], input_shape=[1, 1, 260, 346], logger=quiet, velpred=11,
                           enc_params=ENC, fc_params=FC, form_BEV=1, evs_min_cutoff=0.3, skip_type="none")
    if name == "OrigUNet_w_VITFLY_ViTLSTM":
        return LM.OrigUNet_w_VITFLY_ViTLSTM(num_in_channels=2, num_out_channels=1, num_recurrent=[1, 0], input_shape=[1, 1, 260, 346],
                                            logger=quiet, velpred=0, enc_params=ENC, fc_params=FC, form_BEV=2, evs_min_cutoff=1e-3,
                                            skip_type="interp", is_deployment=False)
    raise KeyError(name)


@pytest.fixture(scope="module")
def manifest(golden_dir):
    return json.load(open(os.path.join(golden_dir, "state_dict_manifest.json")))


NAMES = ["LSTMNetVIT", "ViT", "ConvNet", "LSTMNet", "UNetConvLSTMNet", "OrigUNet_deployed", "OrigUNet_default",
         "OrigUNet_velpred11", "OrigUNet_w_VITFLY_ViTLSTM"]


@pytest.mark.parametrize("name", NAMES)
def test_state_dict_matches_reference_manifest(manifest, name):
    m = build(name)
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert list(ours) == list(manifest[name]), "state_dict key order differs from the reference class"
    assert ours == manifest[name]
    assert all(v.dtype in (torch.float32, torch.int64) for v in m.state_dict().values())
    # checkpoints of the reference load strictly, and state_dict() round-trips them bit for bit
    sd = synth_state_dict(manifest[name], 5)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_reference_attribute_surface():
    m = build("OrigUNet_w_VITFLY_ViTLSTM")
    assert isinstance(m.origunet, LM.OrigUNet) and isinstance(m.vitfly_vitlstm, VM.LSTMNetVIT)
    # run_competition.py:520 reads these
    assert m.vitfly_vitlstm.lstm.num_layers == 3 and m.vitfly_vitlstm.lstm.hidden_size == 128
    assert sum(p.numel() for p in m.parameters()) == 13_420_336       # printed by the reference class
    # sub-model loading with prefixes (learner.py:472-494)
    sd = m.state_dict()
    m.origunet.load_state_dict({k[len("origunet."):]: v for k, v in sd.items() if k.startswith("origunet.")})
    m.vitfly_vitlstm.load_state_dict({k[len("vitfly_vitlstm."):]: v for k, v in sd.items() if k.startswith("vitfly_vitlstm.")})
    m.eval().float()


def test_same_seed_gives_same_initial_weights_as_torch_layers():
    # containers are the same torch layers created in the same order, so the RNG stream matches
    torch.manual_seed(3)
    a = VM.LSTMNetVIT().state_dict()
    torch.manual_seed(3)
    b = VM.LSTMNetVIT().state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    torch.manual_seed(3)
    ref_first = torch.nn.Conv2d(1, 32, kernel_size=7, stride=4, padding=3).weight
    assert torch.equal(a["encoder_blocks.0.patchMerge.cn1.weight"], ref_first)


def test_convlstm_ctor_contract():
    c = ConvLSTM(input_dim=512, hidden_dim=[512], num_layers=1, kernel_size=(1, 1), bias=False, batch_first=True, return_all_layers=False)
    assert list(c.state_dict()) == ["cell_list.0.conv.weight"] and c.state_dict()["cell_list.0.conv.weight"].shape == (2048, 1024, 1, 1)
    with pytest.raises(ValueError):
        ConvLSTM(4, [4], 3, 1)
    h, cc = c._init_hidden(1, (8, 13))[0]
    assert h.shape == (1, 512, 8, 13) and cc.shape == (1, 512, 8, 13)


def test_no_cpu_fallback():
    m = VM.ViT().eval()
    with torch.no_grad(), pytest.raises(_lib.EvflyError):
        m([torch.zeros(1, 1, 60, 90), torch.zeros(1, 1), None])


def test_training_mode_is_refused_whatever_the_grad_mode():
    """The kernels have eval-mode semantics (folded BatchNorm, no Dropout, no power iteration): train() mode would
    diverge from the reference silently, so it raises with or without grad; eval() with grad enabled only warns."""
    m = VM.ViT()      # nn.Module default: training mode
    with pytest.raises(_lib.EvflyError):
        m._check_inference()
    with torch.no_grad(), pytest.raises(_lib.EvflyError):
        m._check_inference()
    from evfly_b200._modbase import PackedModule
    PackedModule._warned_grad = False
    with pytest.warns(UserWarning):
        m.eval()._check_inference()
    with torch.no_grad():
        m._check_inference()
