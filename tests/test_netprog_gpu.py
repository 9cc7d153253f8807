"""The whole-network C entry (danet_net_*, csrc/net.cu) against the Python plan: same kernels, same order, so the
outputs must be IDENTICAL; plus a C host (examples/net_host.c) that never imports Python."""
import os
import subprocess

import numpy as np
import pytest
import torch

from net_common import build, make_image

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_c_executor_equals_python_plan(precision):
    from danet_b200.netprog import NetProgram
    net = build(32, device="cuda:0", conv_algo="auto", precision=precision)
    img = make_image(2, 11).cuda()
    ref = net.infer_net(img)
    ref_vis = [t.clone() for t in ref["visualization"]["iuv_pred"]]
    ref_parts = ref["visualization"]["part_iuv_pred"].clone()
    prog = NetProgram(net.export_program(2), "cuda:0")
    assert prog.batch == 2 and prog.chw == (3, 224, 224) and "para" in prog.names
    for graph in (False, True, True):
        prog.infer(img, graph=graph)
        torch.cuda.synchronize()
        para = prog.output("para").reshape(-1)[:2 * 229].view(2, 229)
        assert torch.equal(para, ref["para"]), (graph, (para - ref["para"]).abs().max().item())
        assert torch.equal(prog.output("centers").reshape(-1)[:96].view(2, 24, 2), ref["stn_kps_pred"])
        for nm, t in zip(("vis_u", "vis_v", "vis_i", "vis_a"), ref_vis):
            assert torch.equal(prog.output(nm), t), nm
        assert torch.equal(prog.output("part_iuv_raw").reshape(ref_parts.shape), ref_parts)
    # another input through the captured graph
    img2 = make_image(2, 12).cuda()
    ref2 = net.infer_net(img2)["para"]
    prog.infer(img2, graph=True)
    torch.cuda.synchronize()
    assert torch.equal(prog.output("para").reshape(-1)[:2 * 229].view(2, 229), ref2)
    # host-buffer convenience entry
    prog.infer_host(img.cpu().numpy())
    assert np.array_equal(prog.read_output("para").reshape(-1)[:2 * 229], ref["para"].cpu().numpy().reshape(-1))
    prog.close()


def test_load_rejects_malformed_programs():
    from danet_b200.netprog import NetProgram
    with pytest.raises(RuntimeError):
        NetProgram(b"not a program" * 20, "cuda:0")
    net = build(32, device="cuda:0", conv_algo="auto")
    blob = net.export_program(1)
    with pytest.raises(RuntimeError):
        NetProgram(blob[:len(blob) // 2], "cuda:0")          # truncated payload
    bad = bytearray(blob)
    bad[8] = 9                                               # version
    with pytest.raises(RuntimeError):
        NetProgram(bytes(bad), "cuda:0")


def test_c_host_without_python(tmp_path):
    exe = os.path.join(ROOT, "examples", "net_host")
    if not os.path.exists(exe):                            # normally built by __graft_entry__.build(); plain gcc, no CUDA headers
        pkg = os.path.join(ROOT, "danet-densepose2smpl_b200")
        subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), exe + ".c", "-o", exe, "-L" + pkg, "-ldanet_b200",
                               "-Wl,-rpath,$ORIGIN/../danet-densepose2smpl_b200"])
    net = build(32, device="cuda:0", conv_algo="auto")
    img = make_image(2, 21)
    ref = net.infer_net(img.cuda())["para"].cpu().numpy()
    prog_path, img_path, out_path = [str(tmp_path / n) for n in ("program.bin", "images.f32", "para.f32")]
    net.export_program(2, prog_path)
    img.numpy().astype(np.float32).tofile(img_path)
    del net
    torch.cuda.empty_cache()
    for mode in ("eager", "graph"):
        r = subprocess.run([exe, prog_path, img_path, out_path, mode], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        para = np.fromfile(out_path, dtype=np.float32)
        assert np.array_equal(para.reshape(-1)[:2 * 229], ref.reshape(-1)), mode
