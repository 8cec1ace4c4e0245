"""The C++ host (`eleven <scene_path> <#samples> <output.bmp>`): loader parity on CPU, end-to-end render on GPU."""
import os
import subprocess

import numpy as np
import pytest

from tfg_pathtracer_b200 import scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tfg-pathtracer_b200", "host", "eleven")


def need_exe():
    assert os.path.exists(EXE), "build the host first: python -c 'import __graft_entry__ as g; g.build()'"


def test_loader_reads_reference_scene_directory(tmp_path):
    need_exe()
    k = S.clock_standin(tex_res=16, xres=48, yres=27, env_size=(64, 32), lights=2)
    k.hdri.xOffset = 0.0
    d = S.write_reference_scene_dir(k, str(tmp_path / "scene"))
    flat = str(tmp_path / "scene.flat")
    out = subprocess.run([EXE, "--dump-flat", d, flat], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    k2 = S.load_flat(flat)
    assert len(k2.tris) == 125281 and (k2.tris["objectID"] == k.tris["objectID"]).all()
    assert np.abs(k2.tris["vertices"] - k.tris["vertices"]).max() < 1e-6          # %.8f text round trip
    assert np.abs(k2.tris["uv"] - k.tris["uv"]).max() < 1e-6
    assert np.abs(k2.tris["normals"] - k.tris["normals"]).max() < 1e-6
    assert (k2.object_material == k.object_material).all()
    assert (k2.materials == k.materials).all()                                      # incl. the reference's texture-id order
    assert all((a.data == b.data).all() and a.format == b.format for a, b in zip(k.textures, k2.textures))
    assert (k2.hdri.data == k.hdri.data).all()                                       # RGBE decode is exact
    assert (k2.lights == k.lights).all() and k2.camera == k.camera


@pytest.mark.parametrize("name", ["cornell", "grid", "clock", "torture"])
def test_loader_tangents_match_the_reference_loader(tmp_path, name):
    """host/tangent_space.cpp against the tangents + handedness THE REFERENCE'S LOADER computed (S/ObjLoader.hpp:167-168: its vendored
    mikktspace per `o` object) for the same scene directory: fixture tests/golden/tangents.npz, generator tests/golden/make_tangents.py
    (runs oracle/_ref/eleven_ref_headless_precise in dump-only mode).  `torture` has UV seams, mirrored islands, degenerate triangles,
    mappings without derivatives, a butterfly edge and corners that keep the default frame."""
    need_exe()
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_tangents as MT
    g = np.load(os.path.join(ROOT, "tests", "golden", "tangents.npz"))
    ours = MT.product_tris(MT.tangent_scenes()[name])
    idx = g[name + "_index"]
    assert len(ours) == int(g[name + "_count"])
    assert (ours["vertices"][idx].view(np.uint32) == g[name + "_vertices"].view(np.uint32)).all()      # same parse, same triangles
    assert np.abs(ours["tangents"][idx] - g[name + "_tangents"]).max() <= 5e-7                            # unit vectors: a few ulp
    assert (ours["tangentsSign"][idx] == g[name + "_sign"]).all()
    if name in ("clock", "torture"):
        assert (g[name + "_sign"] < 0).any() and (g[name + "_sign"] > 0).any()                            # both orientations present


def test_loader_reports_errors(tmp_path):
    need_exe()
    r = subprocess.run([EXE, "--dump-flat", str(tmp_path / "nope"), str(tmp_path / "o.flat")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr
    (tmp_path / "bad").mkdir()
    (tmp_path / "bad" / "scene.json").write_text("{ camera: { xRes: 8 } }")
    r = subprocess.run([EXE, "--dump-flat", str(tmp_path / "bad"), str(tmp_path / "o.flat")], capture_output=True, text=True)
    assert r.returncode != 0 and "resolution" in r.stderr
    assert subprocess.run([EXE], capture_output=True).returncode == 2
    # a texture that cannot be read: reported by the loader's texture threads, the load fails as a whole
    k = S.clock_standin(tex_res=8, xres=16, yres=9, env_size=(16, 8))
    d = S.write_reference_scene_dir(k, str(tmp_path / "scene"))
    victim = sorted(p for p in os.listdir(os.path.join(d, "textures")) if p.endswith(".bmp"))[3]
    os.remove(os.path.join(d, "textures", victim))
    r = subprocess.run([EXE, "--dump-flat", d, str(tmp_path / "o.flat")], capture_output=True, text=True, cwd=d)
    assert r.returncode != 0 and "cannot open" in r.stderr and victim in r.stderr


@pytest.mark.gpu
def test_cli_renders_bmp_like_the_library(tmp_path):
    need_exe()
    from tfg_pathtracer_b200 import renderer as R
    c = S.cornell_box(96, tilt=(3.0, 7.0, 2.0), box_gap=0.002)
    flat = str(tmp_path / "c.flat")
    S.save_flat(c, flat)
    bmp, raw = str(tmp_path / "o.bmp"), str(tmp_path / "o.f32")
    prev, aov = str(tmp_path / "preview.bmp"), str(tmp_path / "aov")
    r = subprocess.run([EXE, flat, "6", bmp, "--mode", "parity", "--raw", raw, "--slice", "2", "--preview", prev, "--aov", aov], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Saved!" in r.stdout and "kPaths/s" in r.stdout
    lib = R.Renderer(**R.PARITY).render_setup(c)
    lib.render_cuda(6)
    film = np.fromfile(raw, np.float32).reshape(96, 96, 4)
    assert (film.view(np.uint32) == lib.film().view(np.uint32)).all()
    img = S.read_bmp(bmp)[::-1]                      # read_bmp returns bottom-up rows; film row 0 is the top row
    assert (img == lib.resolve_rgba8()[..., :3]).all()
    # the last progressive snapshot (after the last 2-sample slice) is the final picture; the AOV files are the first-hit passes
    assert (S.read_bmp(prev) == S.read_bmp(bmp)).all()
    for name, p in (("normal", R.PASS_NORMAL), ("tangent", R.PASS_TANGENT), ("bitangent", R.PASS_BITANGENT)):
        a = np.fromfile("%s_%s.f32" % (aov, name), np.float32).reshape(96, 96, 4)
        assert (a.view(np.uint32) == lib.film(p).view(np.uint32)).all(), name
    # two "GPUs" worth of sample split on one device id is not possible from the CLI; check the fast mode runs
    r = subprocess.run([EXE, flat, "8", bmp], capture_output=True, text=True)
    assert r.returncode == 0 and "pixel-samples/s" in r.stdout
    lib.close()
