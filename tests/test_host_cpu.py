"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares
(no compute calls without a GPU), the drop-in modules reproduce the reference's public
surface (constructor signatures, state_dict schema, parameter counts), and the product
path refuses CPU tensors instead of falling back."""
import ctypes
import inspect
import os

import pytest
import torch

import titanet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from titanet_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 30
    assert os.path.exists(_lib.LIB_PATH), "libtitanet_sm100.so missing: run python -m titanet_b200._build"
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/titanet_b200.h but not exported"
    dll.tn_version.restype = ctypes.c_int
    assert dll.tn_version() >= 100


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "titanet_b200.h")).read()
    for needle in ("src/modules.py", "src/models.py", "src/losses.py", "src/transforms.py"):
        assert needle in text


def test_cpu_tensors_are_refused():
    from titanet_b200 import models, TitanetLibraryError
    net = models.TitaNet.get_titanet(n_mega_blocks=1, model_size="s", dropout=0.0)
    with pytest.raises(TitanetLibraryError):
        net(torch.zeros(2, 80, 50))


@pytest.mark.parametrize("size,blocks,count", [("s", 17, 6151616), ("s", 18, 6428096), ("m", 10, 12904640), ("l", 5, 24690368)])
def test_parameter_counts(size, blocks, count):
    from titanet_b200 import models
    net = models.TitaNet.get_titanet(n_mega_blocks=blocks, model_size=size)
    assert int(net.get_n_params()) == count          # SURVEY.md §4 (measured on the reference)


def test_find_n_mega_blocks_matches_notebook():
    from titanet_b200 import losses, models
    # titanet.ipynb:743-787 -> 18 / 10 / 5 with a CELoss(192, 251) head
    for size, expect in (("m", 10), ("l", 5)):
        assert models.TitaNet.find_n_mega_blocks(192, 80, size, loss_function=losses.CELoss(192, 251)) == expect
    assert models.TitaNet.find_n_mega_blocks(192, 80, "s", loss_function=losses.CELoss(192, 251),
                                             n_mega_blocks_trials=[16, 17, 18, 19]) == 18


@pytest.mark.parametrize("loss", [None, "ce", "arc"])
def test_state_dict_schema(loss):
    from titanet_b200 import losses, models
    lf = {None: None, "ce": losses.CELoss(192, 251), "arc": losses.ArcFaceLoss(192, 251, scale=30, margin=0.2)}[loss]
    net = models.TitaNet.get_titanet(192, 80, 17, "s", loss_function=lf, dropout=0.1)
    sd = net.state_dict()
    schema = O.state_dict_schema(O.TitaNetSpec.named("s", 17), loss, 251)
    assert list(sd.keys()) == list(schema.keys())
    assert all(tuple(sd[k].shape) == tuple(schema[k]) for k in schema)
    assert len(sd) == (643 if loss == "arc" else len(sd))
    net.load_state_dict(O.synth_state_dict(O.TitaNetSpec.named("s", 17), loss, 251), strict=True)


def test_constructor_signatures_match_reference_contract():
    """argument names, order and defaults of the public constructors."""
    from titanet_b200 import losses, models, modules, transforms

    def sig(f):
        return [(p.name, p.default if p.default is not inspect._empty else "<req>") for p in
                list(inspect.signature(f).parameters.values())[1:]]

    assert sig(modules.DepthwiseConv1d.__init__) == [("in_channels", "<req>"), ("out_channels", "<req>"), ("kernel_size", "<req>"),
                                                     ("stride", 1), ("dilation", 1), ("bias", True), ("device", None), ("dtype", None)]
    assert sig(modules.ConvBlock1d.__init__) == [("in_channels", "<req>"), ("out_channels", "<req>"), ("kernel_size", "<req>"),
                                                 ("stride", 1), ("dilation", 1), ("activation", "relu"), ("dropout", 0), ("depthwise", False)]
    assert sig(modules.SqueezeExcitation.__init__) == [("channels", "<req>"), ("reduction", 16)]
    assert sig(models.TitaNet.__init__) == [
        ("n_mels", "<req>"), ("n_mega_blocks", "<req>"), ("n_sub_blocks", "<req>"), ("encoder_hidden_size", "<req>"),
        ("encoder_output_size", "<req>"), ("embedding_size", "<req>"), ("mega_block_kernel_size", "<req>"),
        ("prolog_kernel_size", 3), ("epilog_kernel_size", 1), ("attention_hidden_size", 128), ("se_reduction", 16),
        ("simple_pool", False), ("loss_function", None), ("dropout", 0.5), ("device", "cpu")]
    assert sig(models.MegaBlock.__init__) == [("input_size", "<req>"), ("output_size", "<req>"), ("kernel_size", "<req>"),
                                              ("n_sub_blocks", "<req>"), ("se_reduction", 16), ("dropout", 0.5)]
    assert sig(models.AttentiveStatsPooling.__init__) == [("input_size", "<req>"), ("hidden_size", "<req>"), ("eps", 1e-6)]
    assert sig(losses.ArcFaceLoss.__init__) == [("embedding_size", "<req>"), ("n_classes", "<req>"), ("device", "cpu"),
                                                ("scale", 64), ("margin", 0.5), ("eps", 1e-6)]
    assert sig(losses.AngularMarginLoss.__init__) == [("embedding_size", "<req>"), ("n_classes", "<req>"), ("device", "cpu"),
                                                      ("scale", None), ("m1", 1), ("m2", 0), ("m3", 0), ("eps", 1e-6)]
    assert set(losses.LOSSES) == {"ce", "sphere", "cos", "arc", "ge2e"}
    assert [n for n, _ in sig(transforms.MelSpectrogram.__init__)][:5] == ["sample_rate", "n_fft", "win_length", "hop_length", "n_mels"]
    assert isinstance(losses.ArcFaceLoss(8, 4), losses.MetricLearningLoss)
    with pytest.raises(AssertionError):
        losses.ArcFaceLoss(8, 4, margin=1.5)
    with pytest.raises(AssertionError):
        models.TitaNet.get_titanet(model_size="xl")


def test_mel_host_constants_match_torchaudio_formula():
    from titanet_b200 import TitanetLibraryError, transforms
    assert torch.equal(transforms._htk_filterbank(257, 80, 16000), O.mel_filterbank())
    with pytest.raises(NotImplementedError):
        transforms.MelSpectrogram(16000)          # n_fft=400 default is not a power of two
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80)
    with pytest.raises(TitanetLibraryError):      # no CPU path (this test runs without a GPU)
        mel({"waveform": torch.zeros(1, 16000), "sample_rate": 16000})
    # SpecAugment draws: same generators, same order as the reference (pinned through the oracle's golden test)
    import random
    for seed, frames in ((101, 101), (7, 301)):
        random.seed(seed); torch.manual_seed(seed)
        d = mel.draw_specaugment(frames)
        random.seed(seed); torch.manual_seed(seed)
        rate, fr, fm, tm = O.specaugment_draw(80, frames)
        assert (d.rate, d.frames, d.freq_masks, d.time_masks) == (rate, fr, fm, tm)
        assert 0.95 <= d.rate <= 1.05 and d.frames == transforms.stretched_frames(frames, d.rate)
    off = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    assert off.draw_specaugment(101) is None
    ts = transforms.get_transforms(["chunk"], None)
    assert [type(t).__name__ for t in ts] == ["Resample", "RandomChunk", "MelSpectrogram"]
    assert ts[-1].specaugment_probability == 0.0 and ts[-1].hop_length == 160 and ts[-1].win_length == 400


def test_evaluation_host_logic_buckets_by_length_and_keeps_order():
    """``evaluation.embed_utterances`` (the host half of the batched ``learn.test``): utterances are grouped by frame count,
    each group is forwarded as one batch (never padded together) and the rows come back in the caller's order; Subset /
    indices handling follows src/learn.py:429-434 and src/datasets.py:171.  Device-free: a torch module stands in for the model."""
    from titanet_b200 import evaluation as ev

    class Probe(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))
            self.batches = []

        def forward(self, x):                       # [B, 80, T] -> [B, 3]: (mean, first value, T)
            self.batches.append(tuple(x.shape))
            return torch.stack([x.mean(dim=(1, 2)), x[:, 0, 0], torch.full((x.shape[0],), float(x.shape[2]))], dim=1) * self.w

    g = torch.Generator().manual_seed(0)
    frames = [50, 37, 50, 64, 37, 50, 41]
    specs = [torch.randn(1, 80, t, generator=g) for t in frames]
    model = Probe().train()
    emb = ev.embed_utterances(model, specs, max_batch=2)
    assert not model.training                                            # learn.test puts the model in eval mode (src/learn.py:424)
    assert sorted(model.batches) == sorted([(2, 80, 37), (1, 80, 41), (2, 80, 50), (1, 80, 50), (1, 80, 64)])
    want = torch.cat([Probe()(s) for s in specs]).detach()
    assert torch.allclose(emb, want, atol=1e-6) and emb[:, 2].tolist() == [float(t) for t in frames]
    assert torch.equal(ev.embed_utterances(model, [s[0] for s in specs]), emb)        # [n_mels, T] inputs are accepted
    with pytest.raises(ValueError):
        ev.embed_utterances(model, [torch.randn(2, 80, 10)])
    data = [{"spectrogram": s, "speaker": i % 3} for i, s in enumerate(specs)]
    assert [it["speaker"] for it in ev._dataset_items(data, None)] == [0, 1, 2, 0, 1, 2, 0]
    assert [it["speaker"] for it in ev._dataset_items(data, [4, 0])] == [1, 0]
    assert [it["speaker"] for it in ev._dataset_items(torch.utils.data.Subset(data, [6, 5]), None)] == [0, 2]
    with pytest.raises(Exception):                                        # scoring itself has no CPU path
        ev.cosine_scores(emb)
