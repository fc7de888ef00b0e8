"""GPU parity of the area save-file codec (vx_encode_areas / vx_decode_areas, K8 / K9).
Two LZ4 compressors need not agree on bytes, so parity is defined on what the files decode to:
files written on the GPU are expanded by the oracle's independent LZ4 decoder, and files written by
the oracle's byte-granular compressor (a stand-in for lz4_flex output) are expanded on the GPU."""
import numpy as np
import pytest

import oracle
import voxel_game_b200 as vg

pytestmark = pytest.mark.gpu

SPAWN = 62500


def _files(blob, offsets):
    return [bytes(blob[int(offsets[i]):int(offsets[i + 1])]) for i in range(len(offsets) - 1)]


def _pack(files):
    offsets = np.zeros(len(files) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(f) for f in files])
    blob = np.frombuffer(b"".join(files), np.uint8) if files else np.zeros(0, np.uint8)
    return blob, offsets


def test_encoded_files_decode_to_the_areas(ctx, seed):
    xy = vg.region_xy(SPAWN - 8, SPAWN - 8, 16, 16)
    vox, _ = ctx.generate(xy)
    blob, offsets = ctx.encode_areas(xy)
    assert offsets[0] == 0 and offsets[-1] == len(blob) and np.all(np.diff(offsets.astype(np.int64)) > 0)
    for i, f in enumerate(_files(blob, offsets)):
        assert f[:4] == (32771).to_bytes(4, "little")
        got = oracle.area_file_decode(f)
        assert got is not None, f"area {i}: the oracle's LZ4 decoder rejects the file"
        assert np.array_equal(got, vox[i]), f"area {i} decodes to different voxels"
    assert len(blob) < 0.25 * vox.size  # terrain compresses well even at row granularity


def test_encode_worst_and_crafted_cases(seed):
    """All-literal input (random bytes), every kind of row transition, single-row segments."""
    rng = np.random.default_rng(5)
    n = 6
    vox = np.zeros((n, 32768), np.uint8)
    vox[0] = rng.integers(0, 27, 32768, dtype=np.uint8)                       # no row repeats: one literal run
    vox[1] = 9                                                                 # one row + one giant match
    rows = rng.integers(0, 27, (2048, 16), dtype=np.uint8)
    choice = rng.integers(0, 3, 2048)                                          # 0 new, 1 = previous row, 2 = row one slab up
    for r in range(1, 2048):
        if choice[r] == 1:
            rows[r] = rows[r - 1]
        elif choice[r] == 2 and r >= 16:
            rows[r] = rows[r - 16]
    vox[2] = rows.ravel()
    alt = np.zeros((2048, 16), np.uint8); alt[::2] = 3; alt[1::2] = 7          # period-2 rows: offset-256 matches only
    vox[3] = alt.ravel()
    vox[4] = np.repeat(rng.integers(0, 27, 128, dtype=np.uint8), 256)          # uniform slabs
    vox[5, -16:] = 11                                                          # differs only in the forced-literal last row
    xy = vg.region_xy(100, 200, n, 1)
    with vg.Context(0, seed=seed) as c:
        c.upload(xy, vox, read_heights=False)
        blob, offsets = c.encode_areas(xy)
        files = _files(blob, offsets)
        for i, f in enumerate(files):
            got = oracle.area_file_decode(f)
            assert got is not None and np.array_equal(got, vox[i]), f"case {i}"
        assert len(files[0]) == 4 + 1 + 129 + 32771            # token + 129 length bytes + every byte literal
        assert len(files[1]) < 200 and len(files[0]) <= 33300  # VX_AREA_FILE_MAX_BYTES
        # and back through the GPU decoder
        with vg.Context(0, seed=seed) as d:
            ok, mh = d.decode_areas(xy, blob, offsets)
            assert ok.all()
            v2, mh2 = d.read_areas(xy)
            assert np.array_equal(v2, vox) and np.array_equal(mh, mh2)
            assert np.array_equal(mh, np.stack([oracle.update_all_column_heights(v) for v in vox]))


def test_decode_files_of_another_compressor(seed):
    """Byte-granular LZ4 streams (short offsets, overlapping matches, long length extensions) as a
    general-purpose compressor such as lz4_flex writes them -> resident areas, heights, meshes."""
    n = 6
    ax0, ay0 = SPAWN - 3, SPAWN - 3
    xy = vg.region_xy(ax0, ay0, n, n)
    ovox, omh, _, _ = oracle.generate_areas(seed, xy)
    blob, offsets = _pack([oracle.area_file_encode(v) for v in ovox])
    with vg.Context(0, seed=seed) as c:
        ok, mh = c.decode_areas(xy, blob, offsets)
        assert ok.all() and np.array_equal(mh, omh)
        vox, _ = c.read_areas(xy)
        assert np.array_equal(vox, ovox)
        assert c.resident_count() == n * n
        # the decoded areas are full citizens: meshing them equals meshing generated ones
        inner = vg.region_xy(ax0 + 1, ay0 + 1, n - 2, n - 2)
        info = c.mesh(inner)
        out = c.mesh_read(info)
        with vg.Context(0, seed=seed) as g:
            g.generate(xy, read=False)
            info2 = g.mesh(inner)
            out2 = g.mesh_read(info2)
        assert info == info2
        for a, b in zip(out, out2):
            assert a.tobytes() == b.tobytes()


def test_decode_rejects_bad_files_and_keeps_the_good_ones(seed):
    xy = vg.region_xy(300, 300, 6, 1)
    ovox, omh, _, _ = oracle.generate_areas(seed, xy)
    good = [oracle.area_file_encode(v) for v in ovox]
    bad_id = ovox[3].copy(); bad_id[777] = 27
    payload = bytearray(3 + 32768); payload[0:3] = bytes([251, 1, 0x80])
    files = [good[0],
             good[1][:-9],                                                  # truncated
             bytes([good[2][0] ^ 1]) + good[2][1:],                         # size prefix
             oracle.area_file_encode(bad_id),                               # voxel id 27
             (32771).to_bytes(4, "little") + oracle.lz4_compress_block(bytes(payload)),  # slice length 32769
             good[5]]
    blob, offsets = _pack(files)
    with vg.Context(0, seed=seed) as c:
        ok, mh = c.decode_areas(xy, blob, offsets)
        assert list(ok) == [1, 0, 0, 0, 0, 1]
        assert c.resident_count() == 2
        v, h = c.read_areas(xy[[0, 5]])
        assert np.array_equal(v, ovox[[0, 5]]) and np.array_equal(h, omh[[0, 5]])
        with pytest.raises(vg.VxError) as e:
            c.read_areas(xy[1:2])
        assert e.value.status == -5  # VX_ERR_NOT_LOADED: the caller generates it, like load_blocking
        # world_persistence.rs:66-70: fall back to generation for the rejected ones
        gv, gh = c.generate(xy[1:5])
        assert np.array_equal(gv, ovox[1:5]) and c.resident_count() == 6
        # an empty file and a zero-length batch
        ok2, _ = c.decode_areas(vg.region_xy(900, 900, 1, 1), np.zeros(0, np.uint8), np.zeros(2, np.uint64))
        assert list(ok2) == [0]
        assert c.encode_areas(np.zeros((0, 2), np.uint32), read=False) == 0


def test_encode_needs_resident_areas(seed):
    with vg.Context(0, seed=seed) as c:
        with pytest.raises(vg.VxError) as e:
            c.encode_areas(vg.region_xy(1, 1, 1, 1))
        assert e.value.status == -5


def test_codec_round_trip_full_region(ctx, seed):
    """Config 3 size on the device only: encode 16 384 areas, decode them into a second context,
    compare per-area checksums (size-independent property: decode(encode(x)) == x)."""
    n = 128
    xy = vg.region_xy(SPAWN - n // 2, SPAWN - n // 2, n, n)
    vox, mh = ctx.generate(xy)
    blob, offsets = ctx.encode_areas(xy)
    assert len(blob) < 0.15 * vox.size
    with vg.Context(0, seed=seed) as d:
        ok, mh2 = d.decode_areas(xy, blob, offsets)
        assert ok.all() and np.array_equal(mh2, mh)
        v2, _ = d.read_areas(xy)
        assert np.array_equal(v2, vox)


@pytest.mark.parametrize("world_name,locations", [
    ("test_world_persistence_load_temp_test_world", [(0, 0)]),
    ("test_world_persistence_area_loader_batch_load", [(0, 0), (1, 0), (0, 1), (1, 1)]),
])
def test_reference_persistence_tests(world_name, locations):
    """world_persistence.rs:183-238 (test_world_persistence_load, ..._area_loader_batch_load):
    generate, store, load again, areas equal -- at the reference's own world names and locations
    (the corner of the coordinate space)."""
    seed = vg.hash_world_name(world_name)
    assert seed == oracle.hash_world_name(world_name)
    xy = np.array(locations, np.uint32)
    with vg.Context(0, seed=seed) as c:
        vox, mh = c.generate(xy)
        ovox, omh, _, _ = oracle.generate_areas(seed, xy)
        assert np.array_equal(vox, ovox) and np.array_equal(mh, omh)
        blob, offsets = c.encode_areas(xy)                       # store_all_blocking
        c.unload(xy)
        assert c.resident_count() == 0
        ok, mh2 = c.decode_areas(xy, blob, offsets)              # load_blocking / batch_load
        assert ok.all()
        v2, _ = c.read_areas(xy)
        assert np.array_equal(v2, vox) and np.array_equal(mh2, mh)   # assert_areas_equal


def test_encode_emit_pass_is_speculative_and_backs_off(seed):
    """vx_encode_areas launches its emit pass behind the size pass before the host knows the total, into
    the buffer kept from earlier calls; when the files do not fit, the kernel backs off and the pass is
    launched again.  Small -> large (back-off) -> small (fits) -> large again (fits), each checked."""
    small = vg.region_xy(SPAWN - 2, SPAWN - 2, 3, 3)
    large = vg.region_xy(SPAWN - 10, SPAWN - 10, 20, 20)
    with vg.Context(0, seed=seed) as c:
        vox_s, _ = c.generate(small)
        vox_l, _ = c.generate(large)
        for xy, vox in ((small, vox_s), (large, vox_l), (small, vox_s), (large, vox_l)):
            blob, offsets = c.encode_areas(xy)
            assert offsets[-1] == len(blob)
            for i, f in enumerate(_files(blob, offsets)):
                got = oracle.area_file_decode(f)
                assert got is not None and np.array_equal(got, vox[i])
