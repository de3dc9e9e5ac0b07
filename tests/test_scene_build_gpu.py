"""GPU suite: scene build on the device.  The edge list is built by CUDA kernels between CUB sorts and
scans (redner_b200/csrc/rb_edge_list.cu) and must be the list of the host restatement of collect_edges (src/edge.cpp:233-296), row for
row.  The two secondary-edge trees are built by CUDA kernels
(redner_b200/csrc/rb_edge_tree.cu: Morton codes, radix sort, Karras radix tree, bottom-up bounds, treelet re-optimisation, depth-first
flattening) and must be THE tree the host restatement of EdgeTree::EdgeTree builds (rb_scene_host.hpp, RB_HOST_TREES=1), which the
parity of the hierarchical boundary sampler with the reference depends on: record for record, bit for bit --
child references and every bound; the weighted lengths (a double acos on either side) to one float ulp; the billboard size (a sum
over all edges, reduced in a different order) to 1e-6."""
import numpy as np
import pytest
import torch

import scenes
from redner_b200 import api

pytestmark = pytest.mark.gpu

CASES = [("single_triangle", 32), ("shadow_blocker", 32), ("glossy_room", 32), ("teapot_geometry", 48), ("bunny_box_shifted", 48), ("hires_room", 32), ("random_soup", 32)]


def _trees(rb, dev, scene, res):
    sc = scenes.SCENES[scene](dev, resolution=(res, res))
    args = api.RenderFunction.serialize_scene(sc, 1, 1, sampler_type=rb.SamplerType.sobol, device=dev, backend=rb, use_secondary_edge_sampling=True)
    c = api.RenderFunction._unpack((1, 2), args)
    return c.scene.edge_trees(), c.scene.build_ms(), c.scene.edge_list()


@pytest.mark.parametrize("scene,res", CASES)
def test_gpu_edge_trees_equal_the_host_builder(scene, res, monkeypatch):
    from redner_b200 import redner as rb
    dev = torch.device("cuda:0")
    monkeypatch.setenv("RB_GPU_TREES", "1")  # (small scenes use the host builders by default): edge list and tables on the device
    (rec_g, cs_g, ncs_g, ex_g), ms_g, edges_g = _trees(rb, dev, scene, res)
    monkeypatch.delenv("RB_GPU_TREES")
    monkeypatch.setenv("RB_HOST_TREES", "1")
    (rec_h, cs_h, ncs_h, ex_h), ms_h, edges_h = _trees(rb, dev, scene, res)
    assert edges_g.shape == edges_h.shape and np.array_equal(edges_g, edges_h), "edge lists differ: %s vs %s, first row %s" % (
        edges_g.shape, edges_h.shape, np.argwhere((edges_g != edges_h).any(1))[:1].tolist() if edges_g.shape == edges_h.shape else "-")
    assert rec_g.shape == rec_h.shape and (cs_g, ncs_g) == (cs_h, ncs_h), (rec_g.shape, rec_h.shape, cs_g, cs_h, ncs_g, ncs_h)
    assert abs(ex_g - ex_h) <= 1e-6 * abs(ex_h)
    if rec_g.shape[0] == 0:
        return
    # record = 2 x (pmin[3], pmax[3], dmin[3], dmax[3], wlen, ref) + 4 words of padding
    words = np.ones(32, dtype=bool)
    words[[12, 26]] = False  # wlen of either child
    words[28:] = False
    assert np.array_equal(rec_g[:, words], rec_h[:, words]), "tree topology / bounds differ: %d records" % int((rec_g[:, words] != rec_h[:, words]).any(1).sum())
    wl_g, wl_h = rec_g[:, [12, 26]].view(np.float32), rec_h[:, [12, 26]].view(np.float32)
    assert np.allclose(wl_g, wl_h, rtol=3e-7, atol=0)
    print(scene, "records", rec_g.shape[0], "edge build ms: gpu trees", round(ms_g["edges"], 2), "host trees", round(ms_h["edges"], 2))


def test_edge_list_sides_agree_on_the_full_size_meshes(monkeypatch):
    """The default build (device list from 1024 triangles on) against the host list (RB_HOST_EDGE_LIST=1, tables on the device either
    way) on the C3 / C4 scenes: same list, same trees, and the time of either."""
    from redner_b200 import redner as rb
    dev = torch.device("cuda:0")
    for scene in ("teapot_geometry", "bunny_box_shifted", "hires_room"):
        for k in range(3):  # (the third build of each: pools and caches warm)
            (rec_d, cs_d, ncs_d, _), ms_d, edges_d = _trees(rb, dev, scene, 48)
        monkeypatch.setenv("RB_HOST_EDGE_LIST", "1")
        for k in range(3):
            (rec_h, cs_h, ncs_h, _), ms_h, edges_h = _trees(rb, dev, scene, 48)
        monkeypatch.delenv("RB_HOST_EDGE_LIST")
        assert np.array_equal(edges_d, edges_h) and (cs_d, ncs_d) == (cs_h, ncs_h) and np.array_equal(rec_d[:, :12], rec_h[:, :12])
        print(scene, "edges", edges_d.shape[0], "edge build ms: device list", round(ms_d["edges"], 2), "host list", round(ms_h["edges"], 2),
              "| lights (+ mesh mirror) ms:", round(ms_d["lights"], 2), "vs", round(ms_h["lights"], 2))


def test_batch_of_views_equals_one_scene_per_view():
    """rb_scene_set_camera / api.render_batch: three poses of the C5 teapot through ONE native scene --
    BVH, lights and edge list built once, camera-dependent tables rebuilt on the device per view -- against a full Scene per view:
    same images bit for bit, same summed gradients up to the order of the atomics."""
    from redner_b200 import redner as rb
    import parity_utils as pu
    dev = torch.device("cuda:0")

    def make():
        views = [scenes.teapot_pose(dev, k, num_poses=8, resolution=(48, 48)) for k in (0, 3, 5)]
        for v in views[1:]:  # the views share geometry, materials and lights (the same tensors)
            v.shapes, v.materials, v.area_lights = views[0].shapes, views[0].materials, views[0].area_lights
        return views
    kw = dict(sampler_type=rb.SamplerType.sobol, device=dev, backend=rb)
    views = make()
    imgs = api.render_batch(views, 4, 1, [11, 12, 13], **kw)
    imgs.pow(2).sum().backward()
    g_batch = pu.collect_grads(views[0])
    cam_batch = [v.camera.position.grad.clone() for v in views]
    views = make()
    singles = [api.RenderFunction.apply(11 + k, *api.RenderFunction.serialize_scene(v, 4, 1, **kw)) for k, v in enumerate(views)]
    sum(s.pow(2).sum() for s in singles).backward()
    g_single = pu.collect_grads(views[0])
    for k in range(3):
        assert torch.equal(imgs[k], singles[k]), k
        assert pu.rel_l2(cam_batch[k].numpy(), views[k].camera.position.grad.numpy()) < 1e-4
    for key in g_single:
        if key.startswith("cam."):
            continue
        assert pu.rel_l2(g_batch[key].numpy(), g_single[key].numpy()) < 1e-4, key
