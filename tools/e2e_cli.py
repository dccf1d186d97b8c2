#!/usr/bin/env python
"""End-to-end wall time of the `sabreur` command line on a synthetic file, I/O share visible.

  python tools/e2e_cli.py [--workload cfg3] [--reads 4000000] [--in-format gz|bgzf|bz2|xz|zst|plain] [--out-format zst|gz|none]
                          [--gpus 1] [--inflate-threads N]

Writes the input with the library's own generator, runs the CLI with --timing and prints one JSON line:
wall seconds, reads/s, and the host's split (read+decompress / waiting for the GPU / gather+compress+write).
The GPU kernels are a rounding error here: decompression and compression on the host cores are the wall."""
import argparse
import gzip
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bgzf_blocks(data: bytes, eof: bool = False) -> bytes:
    """gzip members of <= 64 KiB carrying their own size in a 'BC' extra subfield (the BGZF layout of bgzip / htslib)"""
    import struct
    import zlib
    out = bytearray()
    for c in [data[i:i + 65280] for i in range(0, len(data), 65280)] + ([b""] if eof else []):
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        raw = co.compress(c) + co.flush()
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", 18 + len(raw) + 8 - 1)
        out += raw + struct.pack("<II", zlib.crc32(c), len(c))
    return bytes(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--reads", type=int, default=4_000_000)
    # bgzf: blocked gzip as bgzip writes it; bz2: ONE stream of many blocks (what `bzip2` writes); xz: one stream of many
    # blocks (what `xz -T` writes); zst: a frame per 4 MB (what this host, pzstd or `zstd --long`-less chunkers write)
    ap.add_argument("--in-format", default="gz", choices=["gz", "bgzf", "bz2", "xz", "zst", "plain"])
    ap.add_argument("--inflate-threads", type=int, default=0, help="SABREUR_INFLATE_THREADS for the run (0 = the host's default)")
    ap.add_argument("--out-format", default="none", choices=["none", "gz", "zst", "bz2", "xz"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batch-mb", type=int, default=256)
    a = ap.parse_args()
    from sabreur_b200 import synth
    w = synth.WORKLOADS[a.workload]
    tab = synth.table(w)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        paths = []
        for mate in ((0, 1) if w.paired else (0,)):
            ext = {"plain": "", "gz": ".gz", "bgzf": ".gz", "bz2": ".bz2", "xz": ".xz", "zst": ".zst"}[a.in_format]
            p = os.path.join(d, f"in_R{mate + 1}." + ("fa" if w.fmt == 1 else "fq") + ext)
            if a.in_format in ("bz2", "xz", "zst"):  # through a pipe into the format's own compressor
                selftest = os.path.join(ROOT, "sabreur_b200", "bin", "codec_selftest")
                comp = {"bz2": ["bzip2", "-1", "-c"], "xz": ["xz", "-0", "-T0", "--block-size=4MiB", "-c"],
                        "zst": [selftest, "pack", "zst", "1", str(4 << 20)]}[a.in_format]
                with open(p, "wb") as out:
                    pr = subprocess.Popen(comp, stdin=subprocess.PIPE, stdout=out)
                    for lo in range(0, a.reads, 500_000):
                        pr.stdin.write(synth.reads_host(w, tab, lo, min(500_000, a.reads - lo), mate).tobytes())
                    pr.stdin.close()
                    assert pr.wait() == 0
                paths.append(p)
                continue
            opener = (lambda q: gzip.open(q, "wb", compresslevel=1)) if a.in_format == "gz" else (lambda q: open(q, "wb"))
            with opener(p) as fh:
                for lo in range(0, a.reads, 500_000):
                    text = synth.reads_host(w, tab, lo, min(500_000, a.reads - lo), mate).tobytes()
                    fh.write(bgzf_blocks(text) if a.in_format == "bgzf" else text)
                if a.in_format == "bgzf":
                    fh.write(bgzf_blocks(b"", eof=True))
            paths.append(p)
        tsv = os.path.join(d, "bc.tsv")
        with open(tsv, "w") as fh:
            for i, row in enumerate(tab):
                fh.write(f"{bytes(row).decode()}\tbc{i}_R1\tbc{i}_R2\n")
        cmd = [os.path.join(ROOT, "sabreur_b200", "bin", "sabreur"), tsv] + paths + \
              ["-o", os.path.join(d, "out"), "-m", str(w.mismatch), "--gpus", str(a.gpus), "--batch-mb", str(a.batch_mb), "--timing", "-q"]
        if a.out_format != "none":
            cmd += ["-f", a.out_format]
        t0 = time.perf_counter()
        env = dict(os.environ)
        if a.inflate_threads:
            env["SABREUR_INFLATE_THREADS"] = str(a.inflate_threads)
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=d, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            sys.exit(r.stdout + r.stderr)
        m = re.search(r"timing: total ([\d.]+) s \| read\+decompress ([\d.]+) s \| waiting for the GPU ([\d.]+) s \| gather\+compress\+write ([\d.]+) s \| (\d+) bytes in, (\d+) records, (\d+) batches, (\d+) re-submitted", r.stderr)
        total, rd, gw, wr, nbytes, nrec, nb, resub = (float(x) for x in m.groups())
        m2 = re.search(r"timing: process ([\d.]+) s since main\(\) \| GPU setup [^)]*\) ([\d.]+) s", r.stderr)
        process_s, setup_s = (float(x) for x in m2.groups()) if m2 else (None, None)
        in_bytes = sum(os.path.getsize(p) for p in paths)
        out_bytes = sum(os.path.getsize(os.path.join(d, "out", f)) for f in os.listdir(os.path.join(d, "out")))
        print(json.dumps({
            "what": "sabreur CLI end to end (host decompression -> GPU demux -> host compression -> files)",
            "workload": w.name, "reads": int(nrec), "in_format": a.in_format, "out_format": a.out_format, "gpus": a.gpus,
            "wall_s": wall, "reads_per_s": nrec / wall, "input_file_bytes": in_bytes, "text_bytes": int(nbytes), "output_bytes": out_bytes,
            "host_threads": os.cpu_count(), "inflate_threads": a.inflate_threads or "default (min(16, cores))",
            "reader_text_GBps": nbytes / rd / 1e9 if rd > 0 else None,
            "process_s": process_s, "gpu_setup_s": setup_s,
            "pipeline_s": {"total": total, "reader_decompress": rd, "gpu_thread_waiting": gw, "writer_gather_compress_write": wr},
            "batches": int(nb), "resubmitted": int(resub),
            "note": "reader, GPU thread and writer run concurrently: the stages overlap, the slowest one is the wall"}))


if __name__ == "__main__":
    main()
