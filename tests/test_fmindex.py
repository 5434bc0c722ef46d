"""The sort-based FM-index builder against `bwa index` output stored in the golden fixtures."""
import numpy as np
import torch


def test_builder_matches_bwa_index(golden, fm):
    name, z, ix = golden
    got = fm.build_index(z["fwd"])
    assert got.primary == ix.primary
    assert np.array_equal(got.L2, ix.L2)
    assert got.seq_len == ix.seq_len and got.bwt_size == ix.bwt_size
    assert np.array_equal(got.words_numpy(), ix.bwt)


def test_builder_bucketed_path_and_roundtrip(fm, synth, tmp_path):
    ref = synth.make_reference(60000, 5)
    ref[2000:2500] = ref[100:600]
    T = fm.text_from_forward(ref)
    b0, p0 = fm.bwt_string(T, bucket_syms=0)
    for bs in (1, 3):
        b, p = fm.bwt_string(T, bucket_syms=bs)
        assert p == p0 and torch.equal(b, b0)
    ix = fm.pack_bwt(b0, p0)
    path = str(tmp_path / "x.bwt")
    ix.save(path)
    back = fm.BwtIndex.load(path)
    assert back.primary == ix.primary and np.array_equal(back.words_numpy(), ix.words_numpy())
    # LF walk from the primary row must spell the text backwards (definition of the BWT)
    from oracle.binding import Oracle
    o = Oracle(ix)
    n = ix.seq_len
    k = ix.primary
    Tn = T.numpy()
    w = ix.words_numpy()
    for step in range(200):
        # symbol preceding the suffix at row k is T[n-1-step] while walking from the full text
        kk = k - (k >= ix.primary) if k != ix.primary else None
        if step == 0:
            k = 0  # row 0 is "$": its BWT symbol is T[n-1]
        kk = k - (k >= ix.primary)
        c = (int(w[(kk >> 7 << 4) + 8 + ((kk & 127) >> 4)]) >> ((~kk & 15) << 1)) & 3
        assert c == Tn[n - 1 - step]
        k = int(ix.L2[c]) + o.occ4(k)[c]
