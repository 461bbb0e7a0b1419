"""Row F4: gd_mmi_write reproduces the reference's `-d` index dump byte for byte (host logic; the index arrays come from
the oracle here and from the device index in tests/test_gpu_map.py)."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import gdiet_b200 as gd
import maplib
from oraclelib import cpu_has_avx512

pytestmark = pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()),
                                reason="needs oracle/_ref/GDiet_avx_sr (built where /root/reference exists) and AVX-512")


def reference_mmi(contigs, k, w, Z, tmp):
    fa, mmi = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "ref.mmi")
    maplib.write_fasta(fa, contigs)
    subprocess.run([maplib.REF_SR, "-t", "2", "-x", "sr", "-Z", Z, "-W", str(len(Z)), "-k", str(k), "-w", str(w), "-d", mmi, fa],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return open(mmi, "rb").read()


def pack_S(contigs):
    codes = np.concatenate([np.searchsorted(np.frombuffer(b"ACGT", np.uint8), c) for c in contigs]).astype(np.uint32)
    pad = (-len(codes)) % 8
    codes = np.concatenate([codes, np.zeros(pad, np.uint32)]).reshape(-1, 8)
    return (codes << (4 * np.arange(8, dtype=np.uint32))[None, :]).sum(1).astype(np.uint32)


@pytest.mark.parametrize("seed,k,w,Z,lens", [(1, 21, 11, "10", (300000, 200000, 100000)), (2, 15, 10, "10", (250000, 3001)),
                                             (3, 19, 19, "110", (150000,)), (4, 8, 9, "10", (60000, 50000))])
def test_mmi_matches_reference_dump(seed, k, w, Z, lens):
    contigs, _ = maplib.make_dataset(seed=seed, contig_lens=lens, n_reads=1)
    tmp = tempfile.mkdtemp(prefix="gdmmi_")
    want = reference_mmi(contigs, k, w, Z, tmp)
    M = maplib.MapOracle()
    mi = M.index_build(contigs, w, k, Z)
    keys, counts, pos = M.index_arrays(mi)
    M.lib.gdo_index_destroy(mi)
    out = os.path.join(tmp, "ours.mmi")
    gd.mmi_write(out, w, k, ["chr%d" % (i + 1) for i in range(len(contigs))], [len(c) for c in contigs], keys, counts, pos, pack_S(contigs))
    got = open(out, "rb").read()
    assert len(got) == len(want)
    assert hashlib.md5(got).hexdigest() == hashlib.md5(want).hexdigest()
