"""GPU tests of the sort-last composite over peer memory (spv_render_mip_composite): `world` slab contexts in ONE
process on one device stand in for the ranks -- the same kernels, counters and staging layout that run one
process per GPU over NVLink (bench.py --workload slab --composite peer).  The composited image on EVERY rank must
equal the monolithic render bit for bit."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _mono(data, size, M, P, peak, gamma=1.):
    from spimagine_b200 import VolumeRenderer
    r = VolumeRenderer(size)
    r.set_view_copies("primary")  # the slabs hold pairs along z: "bit for bit" is against the single-GPU render through the same (z) copy
    r.set_data(data)
    r.set_modelView(M)
    r.set_projection(P)
    r.render(maxVal=peak, gamma=gamma)
    out = (r.output.copy(), r.output_alpha.copy())
    r.close()
    return out


def _ranks(data, size, world, **kw):
    from spimagine_b200.multigpu import SlabMaxProjector
    rs = []
    for rank in range(world):
        s = SlabMaxProjector(size, rank=rank, world=world, composite="peer", **kw)
        s.set_data(data)  # uploads this rank's slab + halo only
        rs.append(s)
    SlabMaxProjector.connect_local(rs)
    return rs


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("dtype,peak", [(np.uint16, 60000.), (np.float32, 1.)])
def test_peer_composite_is_bit_exact_on_every_rank(world, dtype, peak):
    data = scenes.vol_g(0, dtype, seed=9, shape=(61, 70, 83))
    size = (136, 104)  # 104 rows: bands of 16/20/24/28/52... rows, some cut by the image edge
    rs = _ranks(data, size, world)
    for frame, (theta, gamma) in enumerate([(0.9, 1.), (2.1, 1.), (3.3, .7), (4.0, 1.)]):  # > 2 frames: both parities reused
        M, P = scenes.gui_camera(theta, 2.7)
        want, want_alpha = _mono(data, size, M, P, peak, gamma)
        for s in rs:
            s.set_projection(P)
            s.set_modelView(M)
            s.set_max_val(peak)
            s.set_gamma(gamma)
            s.enqueue_composite()  # nobody blocks before every rank has enqueued
        for s in rs:
            s.collect()
            assert np.array_equal(s.output, want), "world %d rank %d frame %d" % (world, s.rank, frame)
            assert np.array_equal(s.output_alpha, want_alpha)
    for s in rs:
        s.close()


@pytest.mark.parametrize("world,k", [(1, 4), (2, 2), (4, 2), (3, 3)])
def test_several_slabs_per_rank_are_bit_exact(world, k):
    """Serpentine slab assignment (rank r renders slabs r, 2*world-1-r, ...): the helper slabs' partials are
    max-merged on the GPU before the push; the composite still equals the monolithic render bit for bit."""
    data = scenes.vol_g(0, np.uint16, seed=5, shape=(61, 70, 83))
    size = (136, 104)
    rs = _ranks(data, size, world, slabs_per_rank=k)
    assert all(len(s._parts) == k - 1 for s in rs)
    for theta in (0.0, 1.3, 2.9):
        M, P = scenes.gui_camera(theta, 2.7)
        want, want_alpha = _mono(data, size, M, P, 60000.)
        for s in rs:
            s.set_projection(P)
            s.set_modelView(M)
            s.set_max_val(60000.)
            s.enqueue_composite()
        for s in rs:
            s.collect()
            assert np.array_equal(s.output, want) and np.array_equal(s.output_alpha, want_alpha)
    for s in rs:
        s.close()


def test_peer_composite_needs_connect_and_reports_a_missing_peer():
    from spimagine_b200 import _lib
    from spimagine_b200.multigpu import SlabMaxProjector
    data = scenes.vol_g(0, np.uint16, seed=2, shape=(20, 24, 28))
    s = SlabMaxProjector((64, 32), rank=0, world=2, composite="peer")
    s.set_data(data)
    with pytest.raises(RuntimeError):
        s.render()
    s._check(s._lib.spv_comp_init(s._ctx, 0, 2))
    s._connected = True
    with pytest.raises(_lib.SpvError):  # rank 1 was never imported
        s.render()
    with pytest.raises(KeyError):
        SlabMaxProjector((64, 32), rank=0, world=2, composite="nope")
    s.close()


# ----------------------------------------------------------------------------- sort-last iso surface
def _iso_ranks(data, size, world, halo, **kw):
    from spimagine_b200.multigpu import SlabMaxProjector
    rs = []
    for rank in range(world):
        s = SlabMaxProjector(size, rank=rank, world=world, halo=halo, **kw)
        s.set_data(data)
        rs.append(s)
    return rs


def _iso_all(rs, raw_only=False):
    """What SlabMaxProjector._render_isosurface does over NCCL, with the two reductions done here across the
    contexts of one process: MIN over the candidate planes, SUM over the resolved planes."""
    import torch
    for s in rs:
        s.iso_search()
        s.sync()
    k = torch.stack([s.iso_k_tensor() for s in rs]).amin(0)
    for s in rs:
        s.iso_k_tensor().copy_(k)
    torch.cuda.synchronize()
    for s in rs:
        s.iso_resolve()
        s.sync()
    planes = torch.stack([s.iso_planes_tensor() for s in rs]).sum(0)
    for s in rs:
        s.iso_planes_tensor().copy_(planes)
    torch.cuda.synchronize()
    for s in rs:
        s.iso_finish(raw_only)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("dtype,maxval", [(np.uint16, 24000.), (np.float32, .5)])
def test_sort_last_iso_surface_is_bit_exact(world, dtype, maxval):
    """Every rank ends up with the single-GPU iso-surface render: hit mask, depth, normals, occlusion, shading."""
    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.multigpu import iso_halo
    data = scenes.vol_g(0, dtype, seed=7, shape=(64, 72, 80))
    size = (136, 104)
    rs = _iso_ranks(data, size, world, iso_halo(64))
    mono = VolumeRenderer(size)
    mono.set_view_copies("primary")  # (the max projections composited in between: same z copy as the slabs)
    mono.set_data(data)
    for theta, skip, gamma in [(0.4, None, 1.), (1.9, False, 1.), (3.0, None, .8)]:
        M, P = scenes.gui_camera(theta, 3.2)
        for r in rs + [mono]:
            r.set_projection(P)
            r.set_modelView(M)
            r.set_max_val(maxval)
            r.set_gamma(gamma)
            r.set_skipping(skip)
        mono.render(method="iso_surface")
        assert np.isfinite(mono.output_depth).sum() > 500  # there is a surface to talk about
        _iso_all(rs)
        for s in rs:
            assert np.array_equal(s.output_depth, mono.output_depth), (world, s.rank, theta)
            assert np.array_equal(s.output_alpha, mono.output_alpha)
            assert np.array_equal(s.output_normals, mono.output_normals)
            assert np.array_equal(s.output_occlusion, mono.output_occlusion)
            assert np.array_equal(s.output, mono.output)
    for r in rs + [mono]:
        r.close()


def test_sort_last_iso_surface_reports_a_halo_that_is_too_small():
    from spimagine_b200 import _lib
    data = scenes.vol_g(0, np.uint16, seed=7, shape=(64, 72, 80))
    rs = _iso_ranks(data, (96, 80), 4, halo=1)
    M, P = scenes.gui_camera(0.4, 3.2)
    for r in rs:
        r.set_projection(P)
        r.set_modelView(M)
        r.set_max_val(24000.)
    with pytest.raises(_lib.SpvError, match="halo"):
        _iso_all(rs)
    for r in rs:
        r.close()


@pytest.mark.parametrize("sharded", [0, 1])
@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("dtype,maxval", [(np.uint16, 24000.), (np.float32, .5)])
def test_peer_iso_composite_is_bit_exact_on_every_rank(world, dtype, maxval, sharded):
    """spv_render_iso_composite: candidates pushed to the band owners, MIN + redistribution by the owners, finished
    pixels stored into every rank by the rank that owns the crossing -- no reduction on the host side.  Every rank
    ends up with the single-GPU render, frame after frame (the staging alternates by frame parity), also when
    max projections are composited in between on the same contexts."""
    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.multigpu import SlabMaxProjector, iso_halo
    data = scenes.vol_g(0, dtype, seed=7, shape=(64, 72, 80))
    size = (136, 104)
    rs = _iso_ranks(data, size, world, iso_halo(64), composite="peer")
    SlabMaxProjector.connect_local(rs)
    for s in rs:  # knob 12: the screen-space passes on the rank's own band of rows + band gather, or on the whole image
        s._check(s._lib.spv_set_tuning(s._ctx, 12, sharded))
    mono = VolumeRenderer(size)
    mono.set_view_copies("primary")  # (the max projections composited in between: same z copy as the slabs)
    mono.set_data(data)
    for theta, skip, gamma in [(0.4, None, 1.), (1.9, False, 1.), (3.0, None, .8), (4.4, None, 1.)]:
        M, P = scenes.gui_camera(theta, 3.2)
        for r in rs + [mono]:
            r.set_projection(P)
            r.set_modelView(M)
            r.set_max_val(maxval)
            r.set_gamma(gamma)
            r.set_skipping(skip)
        mono.render(method="iso_surface")
        assert np.isfinite(mono.output_depth).sum() > 500
        for s in rs:
            s.enqueue_iso_composite()
        for s in rs:
            s.collect_iso()
        for s in rs:
            assert np.array_equal(s.output_depth, mono.output_depth), (world, s.rank, theta)
            assert np.array_equal(s.output_alpha, mono.output_alpha)
            assert np.array_equal(s.output_normals, mono.output_normals)
            assert np.array_equal(s.output_occlusion, mono.output_occlusion)
            assert np.array_equal(s.output, mono.output)
        if theta > 1.5 and theta < 2.:   # a max projection through the same staging / counters in between
            for r in rs:
                r.set_skipping(False)
            mono.set_skipping(False)
            mono.render(method="max_project")
            for s in rs:
                s.enqueue_composite()
            for s in rs:
                s.collect()
                assert np.array_equal(s.output, mono.output)
    # the iso_surface kernel alone (no post passes)
    mono.render(method="iso_surface_raw")
    for s in rs:
        s.enqueue_iso_composite(raw_only=True)
    for s in rs:
        s.collect_iso()
        assert np.array_equal(s.output_depth, mono.output_depth) and np.array_equal(s.output, mono.output)
        assert np.array_equal(s.output_normals, mono.output_normals)
    for r in rs + [mono]:
        r.close()


def test_peer_iso_composite_reports_a_halo_that_is_too_small():
    from spimagine_b200 import _lib
    from spimagine_b200.multigpu import SlabMaxProjector
    data = scenes.vol_g(0, np.uint16, seed=7, shape=(64, 72, 80))
    rs = _iso_ranks(data, (96, 80), 4, halo=1, composite="peer")
    SlabMaxProjector.connect_local(rs)
    M, P = scenes.gui_camera(0.4, 3.2)
    for r in rs:
        r.set_projection(P)
        r.set_modelView(M)
        r.set_max_val(24000.)
        r.enqueue_iso_composite()
    errors = 0
    for r in rs:
        try:
            r.collect_iso()
        except _lib.SpvError as e:
            assert "halo" in str(e)
            errors += 1
    assert errors >= 1      # the ranks that own a crossing whose taps leave their halo report it
    for r in rs:
        r.close()
