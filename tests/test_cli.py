"""The nextPolish2-compatible command line (nextpolish2_b200/nextPolish2, C++): file readers, option surface, output.
CPU tests use the pass-through path (contigs below -L need no GPU, main.rs:1727-1730); the GPU test polishes a
multi-contig assembly from BAM + FASTA.gz + .yak files and compares the FASTA bytes with the oracle."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import common
import oracle as O
from nextpolish2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nextpolish2_b200", "nextPolish2")


def _run(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, timeout=600, **kw)


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    lens = [60_000, 25_000, 40_000]
    names = ["ctgA", "ctgB", "ctgC"]
    G = synth.genome(909, sum(lens))
    contigs, blobs, haps, off = [], [], [], 0
    for i, L in enumerate(lens):
        A = G[off:off + L].copy()
        off += L
        c = synth.make_contig(910 + i, A, ref_id=i, depth=20, asm_err=4e-4, het=0.001, mean_len=6000, sd_len=800,
                              min_len=2500, threads=2)
        contigs.append(A)
        blobs.append(c["bam"])
        haps += [c["hap1"], c["hap2"]]
    A1 = contigs[1].copy()
    A1[1000:1500] |= 0x20  # a lower-case stretch in the draft
    contigs[1] = A1
    fa = str(d / "asm.fa.gz")
    with gzip.open(fa, "wb") as f:
        for n, s in zip(names, contigs):
            f.write(b">" + n.encode() + b" some description\n")
            b = bytes(s)
            for i in range(0, len(b), 70):
                f.write(b[i:i + 70] + b"\n")
    bam = str(d / "hifi.bam")
    synth.write_bam(bam, names, lens, blobs, level=1)
    yaks = []
    tabs = {}
    for k in (21, 31):
        h, c = synth.make_table(920, k, haps)
        tabs[k] = (h, c)
        p = str(d / ("k%d.yak" % k))
        synth.write_yak(p, k, h, c)
        yaks.append(p)
    return {"dir": d, "fa": fa, "bam": bam, "yaks": yaks, "names": names, "contigs": contigs, "blobs": blobs, "tabs": tabs}


def test_help_and_usage():
    r = _run(["-h"])
    assert r.returncode == 0 and b"Usage: nextPolish2" in r.stderr
    assert _run(["only_one_arg"]).returncode == 2


def test_passthrough_matches_reference_format(files):
    """default -L 1000000: every contig is echoed with start:0 end:len-1, case preserved unless -u"""
    r = _run([files["bam"], files["fa"]] + files["yaks"])
    assert r.returncode == 0, r.stderr
    want = b"".join(O.format_fasta(n, np.arange(len(s), dtype=np.uint32), s) for n, s in zip(files["names"], files["contigs"]))
    assert r.stdout == want
    r = _run(["-u", files["bam"], files["fa"]] + files["yaks"])
    assert r.stdout == want.upper().replace(b"START:", b"start:").replace(b"END:", b"end:").replace(b">CTG", b">ctg")
    out = str(files["dir"] / "o.fa")
    assert _run(["-o", out, files["bam"], files["fa"]] + files["yaks"]).returncode == 0
    assert open(out, "rb").read() == want
    r = _run(["-o", out, files["bam"], files["fa"]] + files["yaks"])  # refuses to overwrite (option.rs:312-316)
    assert r.returncode != 0 and b"already exists" in r.stderr
    r = _run(["--out_pos", files["bam"], files["fa"]] + files["yaks"])
    assert r.stdout.startswith(b"ctgA\t" + bytes(files["contigs"][0][:1]) + b"\t0\nctgA\t")


def test_missing_inputs(files):
    r = _run([files["bam"], str(files["dir"] / "nope.fa")] + files["yaks"])
    assert r.returncode != 0 and b"does not exist" in r.stderr


@pytest.mark.gpu
def test_cli_end_to_end_matches_oracle(files):
    ots = [O.Table.from_arrays(k, *files["tabs"][k]) for k in (21, 31)]
    want = b""
    for n, s, b in zip(files["names"], files["contigs"], files["blobs"]):
        j = O.Job(s, b, ots, O.Opts(min_ctg_len=30_000))
        want += O.format_fasta(n, *j.consensus())
    r = _run(["-L", "30000", "-t", "4", files["bam"], files["fa"]] + files["yaks"][::-1])  # yak order must not matter
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == want
    # the same through zlib on the host threads instead of the device inflate kernel
    r = _run(["-L", "30000", "-t", "4", "--host-inflate", files["bam"], files["fa"]] + files["yaks"])
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == want
    assert want.count(b">") == 3 and bytes(files["contigs"][1]) in want  # ctgB (25 kb) passed through unchanged
    r2 = _run(["-L", "30000", "-g", "1", "--out_pos", files["bam"], files["fa"]] + files["yaks"])
    want_pos = b""
    for n, s, b in zip(files["names"], files["contigs"], files["blobs"]):
        j = O.Job(s, b, ots, O.Opts(min_ctg_len=30_000))
        want_pos += O.format_fasta(n, *j.consensus(), out_pos=True)
    assert r2.returncode == 0 and r2.stdout == want_pos


@pytest.mark.parametrize("level", [0, 1, 6])
def test_bam_reader_returns_every_record_of_a_reference(files, tmp_path, level):
    """BGZF members located through the BAI (first chunk begin .. last chunk end) and inflated in parallel, in place:
    the bytes handed to the polish are exactly the reference's records, for stored and deflated BGZF, any thread count,
    including references whose records start / end in the middle of a member."""
    bam = str(tmp_path / ("l%d.bam" % level))
    synth.write_bam(bam, files["names"], [len(c) for c in files["contigs"]], files["blobs"], level=level)
    for name, blob in zip(files["names"], files["blobs"]):
        for threads in (1, 5):
            r = subprocess.run([CLI, "records", bam, name, str(threads)], capture_output=True, timeout=120)
            assert r.returncode == 0, r.stderr
            assert r.stdout == bytes(blob)
    r = subprocess.run([CLI, "records", bam, "nope"], capture_output=True)
    assert r.returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 1, 6])
def test_bam_reader_with_device_inflate(files, tmp_path, level):
    """the same seam with the members inflated by np2_bgzf_inflate (one warp per member) instead of zlib"""
    bam = str(tmp_path / ("g%d.bam" % level))
    synth.write_bam(bam, files["names"], [len(c) for c in files["contigs"]], files["blobs"], level=level)
    for name, blob in zip(files["names"], files["blobs"]):
        r = subprocess.run([CLI, "records", bam, name, "gpu"], capture_output=True, timeout=120)
        assert r.returncode == 0, r.stderr
        assert r.stdout == bytes(blob)
