"""Randomised end-to-end hunt on the GPU box: the `sabreur` command line (C++ host over the C ABI) on the random cases of
tests/fuzz_gpu.py -- plain / gzip / multi-member gzip input, every output format, 1 MiB batches (guessed cuts, tails,
re-submissions), single- and paired-end -- every output file (decompressed) against the oracle's per-barcode bytes.
Test infrastructure (the oracle is the checker).

    python tests/fuzz_cli.py [--seconds 120] [--seed 1] [--out gpurun_out/fuzz_cli]"""
from __future__ import annotations

import argparse
import bz2
import gzip
import json
import lzma
import os
import random
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_gpu as fz  # noqa: E402
from oracle import oracle as orc  # noqa: E402

SABREUR = os.path.join(ROOT, "sabreur_b200", "bin", "sabreur")
SELFTEST = os.path.join(ROOT, "sabreur_b200", "bin", "codec_selftest")


def plain(path: str) -> bytes:
    data = open(path, "rb").read()
    if not data:
        return b""
    if data[:2] == b"\x1f\x8b":
        return gzip.decompress(data)
    if data[:3] == b"BZh":
        return bz2.decompress(data)
    if data[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(data)
    if data[:4] == b"\x28\xb5\x2f\xfd":
        return subprocess.run([SELFTEST, "cat", path], capture_output=True, check=True).stdout
    return data


def write_input(rng: random.Random, path: str, text: bytes) -> str:
    how = rng.choice(("plain", "plain", "gz", "gz_members", "bz2_blocks", "bz2_streams", "xz_streams", "xz_blocks", "zst_frames"))
    if how == "plain":
        open(path, "wb").write(text)
        return path
    # the splittable formats: many blocks / streams / frames, so that the block-parallel readers (codec_par.cpp) have work
    if how == "bz2_blocks":
        open(path + ".bz2", "wb").write(bz2.compress(text, 1))
        return path + ".bz2"
    if how in ("bz2_streams", "xz_streams"):
        comp = (lambda b: bz2.compress(b, rng.randint(1, 9))) if how == "bz2_streams" else (lambda b: lzma.compress(b, preset=0))
        ext = ".bz2" if how == "bz2_streams" else ".xz"
        cuts = sorted(rng.randrange(len(text) + 1) for _ in range(rng.randint(1, 8)))
        with open(path + ext, "wb") as f:
            for a, b in zip([0] + cuts, cuts + [len(text)]):
                f.write(comp(text[a:b]))
        return path + ext
    if how == "xz_blocks":
        if shutil.which("xz") is None:
            open(path, "wb").write(text)
            return path
        open(path, "wb").write(text)
        subprocess.run(["xz", "-0", "-T2", "--block-size=262144", "-f", path], check=True)
        return path + ".xz"
    if how == "zst_frames":
        with open(path + ".zst", "wb") as f:
            subprocess.run([SELFTEST, "pack", "zst", "1", str(rng.choice((70_001, 1 << 20)))], input=text, stdout=f, check=True)
        return path + ".zst"
    path += ".gz"
    with open(path, "wb") as f:
        if how == "gz":
            f.write(gzip.compress(text, 1))
        else:  # several members, cut at arbitrary bytes (bgzip-like)
            cuts = sorted(rng.randrange(len(text) + 1) for _ in range(rng.randint(1, 6)))
            for a, b in zip([0] + cuts, cuts + [len(text)]):
                f.write(gzip.compress(text[a:b], 1))
    return path


def run_case(seed: int, workdir: str):
    rng = random.Random(seed ^ 0xC11)
    text, bcs, m, _batch, _chunk, _uf, params, starts = fz.make_case(seed)
    if len(text) > 6_000_000:
        return None, params
    err = None
    good = text
    if params["inject"]:  # a violation: the records in front of it are still written (demux.rs:50 / 101, 114)
        try:
            orc.demux(text, bcs, m)
            return "oracle accepted the damaged input (generator bug)", params
        except orc.OracleError as e:
            err = e
        good = text[:starts[err.record]] if err.record < len(starts) else text
    uniq = []
    for b in bcs:  # (a repeated row never wins under first-match; the TSV keeps one file per barcode)
        if b not in uniq:
            uniq.append(b)
    bcs = uniq
    paired = rng.random() < 0.3
    texts = [text]
    goods = [good]
    if paired:  # the mates are independent streams: the reverse file is the forward one with its records reversed
        recs = [text[a:b] for a, b in zip(starts, starts[1:] + [len(text)])]
        texts.append(b"".join(reversed(recs)) if params["tail"] == "nl" and not err else text)
        goods.append(texts[1] if not err else good)
    d = tempfile.mkdtemp(dir=workdir)
    try:
        ext = ".fq" if params["fastq"] else ".fa"
        paths = [write_input(rng, os.path.join(d, f"in_R{i + 1}{ext}"), t) for i, t in enumerate(texts)]
        tsv = os.path.join(d, "bc.tsv")
        with open(tsv, "w") as f:
            for i, b in enumerate(bcs):
                f.write(f"{b.decode()}\tb{i}_R1{ext}" + (f"\tb{i}_R2{ext}" if paired else "") + "\n")
        fmt = rng.choice((None, None, "gz", "zst", "bz2", "xz"))
        # (the host keeps 1/8 of a batch free for carried-over bytes: a record must fit in there)
        batch_mb = rng.choice((4, 8, 256)) if params["shape"] in ("long", "mixed") else rng.choice((1, 1, 2, 256))
        args = [SABREUR, tsv] + paths + ["-o", "out", "-m", str(m), "--batch-mb", str(batch_mb), "-q"]
        if fmt:
            args += ["-f", fmt, "-l", str(rng.randint(1, 4))]
        if paired and rng.random() < 0.3:
            args.append("--sequential-mates")
        args += os.environ.get("SABREUR_EXTRA_ARGS", "").split()  # e.g. "--gpus 2" on a multi-GPU box
        params.update(paired=paired, out_fmt=fmt, inputs=[os.path.basename(p) for p in paths], n_bc=len(bcs))
        r = subprocess.run(args, capture_output=True, text=True, cwd=d, timeout=300)
        # SE: `record?` surfaces the parse error (exit 1); PE: `while let Some(Ok(..))` just stops that mate; a read
        # shorter than the barcode is the slice panic (exit 101) in both
        want_rc = 0 if not err else (101 if err.code == orc.E_SHORT_READ else (0 if paired else 1))
        params["want_rc"] = want_rc
        if r.returncode != want_rc:
            return f"exit {r.returncode}, expected {want_rc}: {r.stderr[-400:]}", params
        out = os.path.join(d, "out")
        files = {}
        for name in os.listdir(out):
            stem = name
            for e in (".gz", ".zst", ".bz2", ".xz"):
                if stem.endswith(e):
                    stem = stem[:-len(e)]
            files[stem] = os.path.join(out, name)
        for mate, t in enumerate(goods):
            if want_rc == 101 and paired and mate == 1:
                # the forward pass panicked: the reference never reaches the reverse loop (demux.rs:101-123), and the
                # host takes back whatever its side-by-side reverse pass had written
                for stem, path_ in files.items():
                    if "_R2" in stem and os.path.getsize(path_) != 0:
                        return f"{stem} is not empty although the forward pass panicked", params
                break
            exp = orc.demux_outputs(t, bcs, m) if t else [b""] * (len(bcs) + 1)
            for i in range(len(bcs)):
                stem = f"b{i}_R{mate + 1}{ext}"
                if stem not in files:
                    return f"missing {stem}", params
                if plain(files[stem]) != exp[i]:
                    return f"{stem} differs from the oracle's bytes", params
            unk = ("unknown_R%d.fa" % (mate + 1)) if paired else "unknown.fa"
            got = plain(files[unk]) if unk in files else b""
            if got != exp[-1]:
                return f"{unk} differs from the oracle's bytes", params
            if unk in files and not exp[-1] and r.returncode == 0:
                return f"{unk} exists but should have been deleted", params
        return "", params
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/fuzz_cli")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if not os.path.exists(SABREUR):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "sabreur_b200", "csrc", "host")])
    work = tempfile.mkdtemp(prefix="fuzzcli")
    t0 = time.time()
    n = ran = fails = 0
    while time.time() - t0 < a.seconds:
        seed = a.seed * 7_000_003 + n
        n += 1
        try:
            why, params = run_case(seed, work)
        except Exception as e:  # noqa: BLE001
            why, params = f"exception: {e!r}", dict(seed=seed)
        if why is None:
            continue
        ran += 1
        if why:
            fails += 1
            params["why"] = why
            print("FAIL", json.dumps(params)[:700], flush=True)
            with open(os.path.join(a.out, f"case_{seed}.json"), "w") as f:
                json.dump(params, f)
            if fails >= 5:
                break
    shutil.rmtree(work, ignore_errors=True)
    print(f"fuzz_cli: {ran} cases, {fails} failures, {time.time() - t0:.0f} s", flush=True)
    with open(os.path.join(a.out, "summary.json"), "w") as f:
        json.dump(dict(cases=ran, failures=fails, seconds=time.time() - t0, seed=a.seed), f)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
