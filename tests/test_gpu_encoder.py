"""K1/K2 parity (bit-exact, integer work): GPU encoder vs the oracle's CollectFeatures / TransformFeature."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    from unrealgo_b200 import DlTFNetworkEvaluator
    e = DlTFNetworkEvaluator("", max_batch=4096)
    yield e
    e.close()


def test_encoder_planes_bit_exact(ev, positions):
    packed, planes = positions
    got = ev.encode_planes(packed)
    assert got.dtype == np.int8 and got.shape == planes.shape
    assert np.array_equal(got, planes)


def test_encoder_nhwc_is_transform_feature(ev, positions):
    import oracle
    packed, planes = positions
    want = oracle.transform_feature(planes, data_format=0, as_bool=True)  # bool[n,19,19,17] handed to TF
    assert np.array_equal(ev.encode_nhwc(packed), want)
    assert np.array_equal(ev.TransformFeature(planes), want)


def test_transform_feature_bool_semantics(ev):
    import oracle
    rng = np.random.default_rng(5)
    planes = rng.integers(-3, 4, size=(9, 17, 19, 19)).astype(np.int8)  # (bool)char: any non-zero is 1
    assert np.array_equal(ev.TransformFeature(planes), oracle.transform_feature(planes, 0, True))


def test_encoder_many_positions(ev):
    import oracle
    total = 0
    for seed in (77, 78, 79):  # 10 500 seeded positions, in evaluator-sized chunks
        packed, planes = oracle.synthetic_positions(3500, seed=seed)
        assert np.array_equal(ev.encode_planes(packed), planes)
        assert np.array_equal(ev.encode_nhwc(packed), oracle.transform_feature(planes, 0, True))
        total += len(packed)
    assert total >= 10000


def test_encoder_matches_the_reference_boards_planes(ev):
    """planes written by the REFERENCE's own board driven like CollectFeatures (tests/golden/ref_board.npz, generator
    tests/golden/make_ref_board_golden.py): the games are replayed on the host board the self-play driver uses
    (csrc/goboard.h), its packed history goes through the CUDA encoder, and the result must be those planes"""
    import os
    from unrealgo_b200 import ProductBoard
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_board.npz"))
    packed, want = [], []
    for g in range(int(G["op_game"].max()) + 1):
        sel = np.nonzero(G["op_game"] == g)[0]
        ops, cols = G["ops"][sel].tolist(), G["op_color"][sel].tolist()
        if -3 in ops:
            cut = ops.index(-3)
            ops, cols = ops[:cut], cols[:cut]
        if -2 in ops or cols != [k & 1 for k in range(len(ops))]:
            continue
        samples = {int(G["s_after"][i]): i for i in np.nonzero(G["s_game"] == g)[0]}
        pb = ProductBoard()
        for n, op in enumerate([None] + ops):
            if op is not None:
                pb.play(op)
            if n in samples and G["s_has_planes"][samples[n]]:
                packed.append(pb.pack())
                want.append(np.unpackbits(G["s_planes"][samples[n]])[:17 * 361].reshape(17, 19, 19).astype(np.int8))
    assert len(packed) > 350
    assert np.array_equal(ev.encode_planes(np.stack(packed)), np.stack(want))


def test_encoder_matches_the_reference_collect_features_template(ev):
    """tests/golden/ref_collect_features.npz holds the output of the reference's own CollectFeatures template
    (`ref_search features`, see tests/golden/make_ref_board_golden.py) after every move of 21 games — 2966 positions.
    Packed histories of the same games from the self-play driver's board, through the CUDA encoder, must equal it."""
    import os
    from unrealgo_b200 import ProductBoard
    F = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_collect_features.npz"))
    want = np.unpackbits(F["planes"], axis=1)[:, :17 * 361].reshape(-1, 17, 19, 19).astype(np.int8)
    packed, at = [], 0
    for n in F["game_len"].tolist():
        pb = ProductBoard()
        packed.append(pb.pack())
        for mv in F["moves"][at:at + n].tolist():
            pb.play(mv)
            packed.append(pb.pack())
        at += n
    assert len(packed) == len(want) == 2966
    assert np.array_equal(ev.encode_planes(np.stack(packed)), want)
