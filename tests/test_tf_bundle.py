"""The TF tensor-bundle writer/reader against the reference's archived checkpoints
(format fixtures copied to tests/golden/)."""
import os

import numpy as np

from graphembeddings_b200 import tf_bundle as T


def test_crc32c_check_value():
    assert T.crc32c(b"123456789") == 0xE3069283
    a = np.arange(100000, dtype=np.float32)
    assert T.crc32c(a) == T._crc32c_py(a.tobytes())          # native helper == pure python


def test_reader_parses_archived_index_files(golden_dir):
    e = dict(T.read_index(os.path.join(golden_dir, "ref_holE-20170712_model.ckpt.index")))
    assert list(e) == [b"", b"embeddings"]
    ent = T.decode_entry(e[b"embeddings"])
    assert ent["dtype"] == T.DT_FLOAT and ent["shape"] == [35910, 128]
    assert ent["size"] == 35910 * 128 * 4 and ent["offset"] == 0
    e2 = T.read_index(os.path.join(golden_dir, "ref_holE-20170724_model.ckpt.index"))
    keys = [k.decode() for k, _ in e2]
    assert keys == ["", "batch/Variable", "embeddings", "id_to_type-keys", "id_to_type-values",
                    "type_to_ids-keys", "type_to_ids-values"]
    emb = T.decode_entry(dict(e2)[b"embeddings"])
    assert emb["shape"] == [1134637, 128] and emb["offset"] == 4 and emb["size"] == 1134637 * 128 * 4
    tti = T.decode_entry(dict(e2)[b"type_to_ids-values"])
    assert tti["dtype"] == T.DT_INT64 and tti["shape"] == [12, 100000]     # 12 types x padded_size


def test_writer_reproduces_archived_index_files_byte_for_byte(golden_dir, tmp_path):
    for name in ("ref_holE-20170712_model.ckpt.index", "ref_holE-20170724_model.ckpt.index"):
        src = os.path.join(golden_dir, name)
        entries = T.read_index(src)
        # re-encode every record from its decoded fields, then rebuild the table
        rebuilt = []
        for k, v in entries:
            if k == b"":
                rebuilt.append((k, T.encode_header()))
            else:
                e = T.decode_entry(v)
                rebuilt.append((k, T.encode_entry(e["dtype"], e["shape"], e["offset"], e["size"],
                                                  e["crc32c"], e["shard_id"])))
        out = os.path.join(tmp_path, name)
        T.write_index(out, rebuilt)
        assert open(out, "rb").read() == open(src, "rb").read()


def test_bundle_roundtrip_and_text_files(golden_dir, tmp_path):
    rng = np.random.default_rng(0)
    E = rng.standard_normal((300, 150)).astype(np.float32)
    prefix = os.path.join(tmp_path, "model.ckpt")
    T.save_bundle(prefix, {"embeddings": E, "batch/Variable": np.array(7, dtype=np.int32)})
    got = T.load_bundle(prefix)
    assert np.array_equal(got["embeddings"], E) and int(got["batch/Variable"]) == 7
    assert os.path.getsize(prefix + ".data-00000-of-00001") == 4 + E.nbytes
    # corruption is detected
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(100); f.write(b"\xff")
    try:
        T.load_bundle(prefix)
        assert False, "checksum mismatch not detected"
    except ValueError:
        pass
    T.write_checkpoint_state(str(tmp_path))
    ref = open(os.path.join(golden_dir, "ref_holE-20170712_checkpoint")).read().splitlines()
    mine = open(os.path.join(tmp_path, "checkpoint")).read().splitlines()
    assert mine[0] == ref[0] and mine[-1] == ref[-1]
    T.write_projector_config(str(tmp_path))
    assert open(os.path.join(tmp_path, "projector_config.pbtxt")).read() == \
        open(os.path.join(golden_dir, "ref_holE-20170712_projector_config.pbtxt")).read()
