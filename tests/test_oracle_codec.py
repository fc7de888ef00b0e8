"""Oracle for the area save-file format (oracle/codec.c): LZ4 block format, bincode framing of
AreaDTO (generic_persistence.rs:17-103, world_persistence.rs:27-71, area.rs:204-227)."""
import numpy as np

import oracle


def test_lz4_block_known_streams():
    # token 0x44: 4 literals + match length 4+4, offset 4 (overlapping copy); closing token 0x50: 5 literals
    blk = bytes([0x44]) + b"abcd" + bytes([4, 0]) + bytes([0x50]) + b"12345"
    assert oracle.lz4_decompress_block(blk, 64) == b"abcdabcdabcd12345"
    # run-length case: one literal, offset 1, length 4 + 15 + 255 + 3 = 277
    blk = bytes([0x1F]) + b"x" + bytes([1, 0]) + bytes([255, 3]) + bytes([0x50]) + b"vwxyz"
    assert oracle.lz4_decompress_block(blk, 1024) == b"x" * 278 + b"vwxyz"
    # literal length extension: 15 + 255 + 0 = 270 literals
    lit = bytes(range(256)) + bytes(range(14))
    blk = bytes([0xF0, 255, 0]) + lit
    assert oracle.lz4_decompress_block(blk, 1024) == lit


def test_lz4_block_rejects_malformed_streams():
    assert oracle.lz4_decompress_block(bytes([0x44]) + b"abcd" + bytes([5, 0]), 64) is None   # offset beyond start
    assert oracle.lz4_decompress_block(bytes([0x44]) + b"abcd" + bytes([0, 0]), 64) is None   # offset 0
    assert oracle.lz4_decompress_block(bytes([0x40]) + b"abc", 64) is None                    # literals cut short
    assert oracle.lz4_decompress_block(bytes([0x44]) + b"abcd" + bytes([4]), 64) is None      # offset cut short
    assert oracle.lz4_decompress_block(bytes([0x4F]) + b"abcd" + bytes([4, 0, 255]), 64) is None  # length ext cut
    assert oracle.lz4_decompress_block(bytes([0x44]) + b"abcd" + bytes([4, 0]), 10) is None   # output too small


def test_compressor_round_trips_and_respects_end_rules():
    rng = np.random.default_rng(1)
    for data in (bytes(1000), bytes(rng.integers(0, 3, 5000, dtype=np.uint8)), bytes(rng.integers(0, 256, 300, dtype=np.uint8)),
                 b"abc" * 500, b"0123456789abc"):
        c = oracle.lz4_compress_block(data)
        assert oracle.lz4_decompress_block(c, len(data)) == data
    assert len(oracle.lz4_compress_block(bytes(32771))) < 200


def test_area_file_layout(seed):
    vox, _, _, _ = oracle.generate_areas(seed, oracle.region_xy(62500, 62500, 2, 1))
    f = oracle.area_file_encode(vox[0])
    assert f[:4] == (32771).to_bytes(4, "little")            # compress_prepend_size
    payload = oracle.lz4_decompress_block(f[4:], 32771)
    assert payload[:3] == bytes([251, 0x00, 0x80])            # bincode varint: 251 marker + u16 LE 32768
    assert payload[3:] == vox[0].tobytes()                    # one variant byte per voxel
    assert len(f) < 4000
    assert np.array_equal(oracle.area_file_decode(f), vox[0])


def test_area_file_decode_rejects_what_the_reference_rejects(seed):
    vox, _, _, _ = oracle.generate_areas(seed, oracle.region_xy(62500, 62500, 1, 1))
    f = bytearray(oracle.area_file_encode(vox[0]))
    assert oracle.area_file_decode(bytes(f[:-7])) is None                         # truncated
    assert oracle.area_file_decode(bytes(f[:3])) is None
    bad = bytearray(f); bad[0] ^= 1
    assert oracle.area_file_decode(bytes(bad)) is None                            # size prefix
    v = vox[0].copy(); v[12345] = 27                                              # not a Voxel variant
    assert oracle.area_file_decode(oracle.area_file_encode(v)) is None
    payload = bytearray(3 + 32768); payload[0:3] = bytes([251, 1, 0x80])          # slice length 32769
    blk = (32771).to_bytes(4, "little") + oracle.lz4_compress_block(bytes(payload))
    assert oracle.area_file_decode(blk) is None
