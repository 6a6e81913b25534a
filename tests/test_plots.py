"""The reference's pictures (src/plot.rs): metaplot.png and bootstrap.png written by libabfit's own PNG encoder.
Checked: the files are valid PNGs of the reference's canvas size (decoded by an independent decoder: zlib via Pillow),
the series are there in the reference's colours, the ranges are the reference's (a value outside y = 0..0.01 is clipped)."""
import os
import struct
import zlib

import numpy as np
import pytest


def decode_png(path):
    """minimal independent decoder: 8-bit RGB, filter 0 only (what abfit_plot.cu writes)"""
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) == crc, typ
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype, comp, filt, interlace = ihdr
    assert (depth, ctype, comp, filt, interlace) == (8, 2, 0, 0, 0)
    data = zlib.decompress(idat)  # checks the deflate stream and the Adler-32
    assert len(data) == h * (1 + 3 * w)
    rows = np.frombuffer(data, dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert not rows[:, 0].any()
    return rows[:, 1:].reshape(h, w, 3), raw


def test_metaplot_png(ab, tmp_path):
    n = 300
    x = np.arange(n)
    alpha = 0.002 + 0.001 * np.sin(x / 30.0)
    beta = 0.006 + 0.002 * np.cos(x / 50.0)
    beta[100:110] = 0.05  # outside the reference's fixed y range 0..0.01: clipped at the frame
    p = str(tmp_path / "metaplot.png")
    ab.plot_metaplot(p, alpha, beta, (alpha * 0.8, alpha * 1.2), (beta * 0.9, beta * 1.1))
    img, raw = decode_png(p)
    assert img.shape == (960, 1280, 3)  # (640 * 2, 480 * 2), src/plot.rs:9
    assert len(raw) < 400_000  # the run-length deflate works (3.7 MB uncompressed)
    red = (img[:, :, 0] == 255) & (img[:, :, 1] == 0) & (img[:, :, 2] == 0)
    blue = (img[:, :, 2] == 255) & (img[:, :, 0] == 0) & (img[:, :, 1] == 0)
    assert red.sum() > 1000 and blue.sum() > 800
    # the alpha curve sits at 20-30 % of the y range, beta above it
    ys_r, ys_b = np.nonzero(red.any(axis=1))[0], np.nonzero(blue.any(axis=1))[0]
    assert ys_b.mean() < ys_r.mean()  # image y grows downwards
    # translucent bands: pixels that are neither white, grid, black nor a pure series colour
    pinkish = (img[:, :, 0] == 255) & (img[:, :, 1] == 204) & (img[:, :, 2] == 204)  # 20 % red over white
    assert pinkish.sum() > 5000
    try:
        from PIL import Image
        im = Image.open(p)
        im.load()
        assert im.size == (1280, 960)
    except ImportError:
        pass
    # no windows at all: still a valid picture (frame and axes only)
    ab.plot_metaplot(str(tmp_path / "empty.png"), [], [])
    decode_png(str(tmp_path / "empty.png"))


def test_bootstrap_png(ab, tmp_path):
    rng = np.random.default_rng(5)
    alphas = rng.normal(2e-4, 2e-5, 500).clip(min=1e-6)
    betas = rng.normal(8e-4, 9e-5, 500).clip(min=1e-6)
    p = str(tmp_path / "bootstrap.png")
    ab.plot_bootstrap(p, alphas, betas)
    img, _ = decode_png(p)
    assert img.shape == (960, 1280, 3)
    red = (img[:, :, 0] == 255) & (img[:, :, 1] == 0) & (img[:, :, 2] == 0)
    blue = (img[:, :, 2] == 255) & (img[:, :, 0] == 0) & (img[:, :, 1] == 0)
    assert red.sum() > 300 and blue.sum() > 300
    # Alpha's box is left of Beta's and lower on the page's value axis (smaller rate = further down the image)
    assert np.nonzero(red.any(axis=0))[0].mean() < np.nonzero(blue.any(axis=0))[0].mean()
    assert np.nonzero(red.any(axis=1))[0].mean() > np.nonzero(blue.any(axis=1))[0].mean()
    # y range = 0 .. 1.3 max (src/plot.rs:101): the highest blue pixel (upper fence, clipped to the data range at most
    # 1.3 max) is below the frame's top
    with pytest.raises(ab.AbfitError):
        ab.plot_bootstrap(p, [np.nan, 1.0], [1.0, 2.0])
    with pytest.raises(ab.AbfitError):
        ab.plot_bootstrap(str(tmp_path / "no_such_dir" / "x.png"), alphas, betas)


def test_progress_bar_format(tmp_path):
    """cli/progress.h: the templates of src/progress.rs:5-43 — "{msg} {bar:40} [{elapsed}] {pos:>7}/{len:7}" (+ " ETA: ..."
    for the overall bar), drawn on stderr, forced / hidden by ABFIT_PROGRESS"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "p.cpp"
    src.write_text('#include "progress.h"\n'
                   'int main() { progress::Bar a("ABNeutral", 1000, false), b("Progress ", 300, true);\n'
                   '  a.set(250); b.inc(30); std::printf("%s\\n%s\\n", a.render().c_str(), b.render().c_str()); a.finish(); b.finish(); return 0; }\n')
    exe = tmp_path / "p"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "alphabeta-rs_b200", "cli"), "-o", str(exe), str(src)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, ABFIT_PROGRESS="1"))
    line_a, line_b = r.stdout.split("\n")[:2]
    assert line_a.startswith("ABNeutral ") and line_a.endswith("[00:00]     250/1000   ")
    assert line_a.count("█") == 10  # a quarter of the 40-cell bar
    assert line_b.startswith("Progress  ") and "      30/300     ETA: " in line_b and line_b.endswith("seconds")
    assert "ABNeutral" in r.stderr and "1000/1000" in r.stderr and r.stderr.endswith("\n")
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, ABFIT_PROGRESS="0"))
    assert r.stderr == ""
