"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): the drop-in claim of INTEGRATION.md §1
executed with the reference's OWN code.

`im2mesh.metaavatar_render.config.get_model` (config.py:147-302) is run on the reference's shipped YAML configs twice: unmodified,
and with the two names `IDHRNetwork` / `BodyRayTracing` in `im2mesh/metaavatar_render/models/__init__.py` pointing at
`arah_release_b200.renderer` (the three-line change).  The second run goes through the reference's own test-time path
(`mode='test'`: constructor call of `MetaAvatarRender` with the YAML's arguments, checkpoint load at :291-300) and must give a model
with the same state_dict keys, load the first model's weights strictly, keep the tracer attributes the reference reads, and —
since there is no CPU path — refuse to render CPU tensors.  Stubs (oracle/ref_harness.py) only stand in for absent packages."""
import os
import tempfile
import types

import pytest
import torch

REF = os.environ.get('ARAH_REFERENCE_ROOT', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'im2mesh')), reason='needs the reference tree (build container only)')


def _cfg(name):
    import yaml

    def merge(a, b):
        for k, v in b.items():
            if isinstance(v, dict) and isinstance(a.get(k), dict):
                merge(a[k], v)
            else:
                a[k] = v
        return a
    with open(os.path.join(REF, 'configs', 'default.yaml')) as f:
        base = yaml.safe_load(f)
    with open(os.path.join(REF, 'configs', name)) as f:
        cfg = merge(base, yaml.safe_load(f))
    cfg['model']['train_smpl'] = False            # optimised SMPL parameters need the dataset's files; not part of the renderer
    return cfg


@pytest.mark.parametrize('name', ['arah-zju/ZJUMOCAP-377_4gpus.yaml', 'arah-zju/ZJUMOCAP-377-mono_4gpus.yaml', 'arah-h36m/H36M_S9_4gpus.yaml'])
def test_reference_get_model_builds_with_our_classes(name, monkeypatch):
    from oracle import ref_harness as rh
    rh.install()
    import im2mesh.metaavatar_render  # noqa: F401
    from im2mesh.metaavatar_render import config as mcfg, models as mmodels
    from arah_release_b200 import _lib, renderer as ours
    cfg = _cfg(name)
    dataset = types.SimpleNamespace(data=[{'cam_idx': 0, 'frame_idx': i} for i in range(3)], cam_names=[], cameras={})
    ref_model = mcfg.get_model(cfg, mode='val', dataset=dataset)                       # unmodified reference
    assert type(ref_model.idhr_network).__module__.startswith('im2mesh.')
    sd = ref_model.state_dict()
    with tempfile.TemporaryDirectory() as d:
        ckpt = os.path.join(d, 'model.ckpt')
        torch.save({'state_dict': {'model.' + k: v for k, v in sd.items()}}, ckpt)     # Lightning layout (config.py:294-300)
        monkeypatch.setattr(mmodels, 'IDHRNetwork', ours.IDHRNetwork)                   # the change of INTEGRATION.md §1
        monkeypatch.setattr(mmodels, 'BodyRayTracing', ours.BodyRayTracing)
        model = mcfg.get_model(cfg, mode='test', checkpoint_path=ckpt)
    assert isinstance(model.idhr_network, ours.IDHRNetwork) and isinstance(model.idhr_network.ray_tracer, ours.BodyRayTracing)
    sd2 = model.state_dict()
    assert set(sd2) == set(sd)
    model.load_state_dict(sd, strict=True)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)                                 # the reference's own loader filled our modules
    # aliases of models/__init__.py:66-76 are the same objects
    assert model.idhr_network.rendering_network is model.color_decoder and model.idhr_network.skinning_model is model.skinning_model
    assert model.idhr_network.deviation_network is model.deviation_decoder
    m = cfg['model']
    tr = model.idhr_network.ray_tracer
    assert (tr.n_steps, tr.near_surface_vol_samples, tr.far_surface_vol_samples) == (m['n_steps'], m['near_surface_samples'], m['far_surface_samples'])
    assert model.idhr_network.cano_view_dirs == m['cano_view_dirs']
    # no CPU path behind the reference's module tree either
    from arah_release_b200 import synthetic as syn
    from tools import ref_layout as rl
    fr = syn.make_frame(8, 8, seed=0)
    with pytest.raises(_lib.ArahError):
        model.idhr_network.eval()(rl.inputs_from_frame(fr, rl.sdf_network_from_frame(fr, 'cpu'), 'cpu'))


def test_constructor_and_forward_signatures_accept_every_reference_argument():
    """Every parameter of the reference's constructors / forwards exists in ours under the same name, in the same position and with
    the same default (ours may append optional extras such as shade_mode)."""
    import inspect
    from oracle import ref_harness as rh
    rh.install()
    import im2mesh.metaavatar_render  # noqa: F401
    from im2mesh.metaavatar_render.renderer.implicit_differentiable_renderer import IDHRNetwork as RefNet
    from im2mesh.metaavatar_render.renderer.ray_tracing import BodyRayTracing as RefTracer
    from im2mesh.metaavatar_render.renderer.loss import IDHRLoss as RefLoss
    from arah_release_b200.loss import IDHRLoss
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    for ref, ours, methods in ((RefNet, IDHRNetwork, ('__init__', 'forward')), (RefTracer, BodyRayTracing, ('__init__', 'forward')),
                               (RefLoss, IDHRLoss, ('__init__', 'forward'))):
        for meth in methods:
            pr = list(inspect.signature(getattr(ref, meth)).parameters.values())
            po = list(inspect.signature(getattr(ours, meth)).parameters.values())
            kinds = {p.kind for p in po}
            for i, p in enumerate(pr):
                if p.kind in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL):
                    continue
                match = [q for q in po if q.name == p.name]
                assert match or inspect.Parameter.VAR_KEYWORD in kinds, f'{ours.__name__}.{meth} lacks the reference argument {p.name!r}'
                if match:
                    assert po.index(match[0]) == i, f'{ours.__name__}.{meth}: {p.name!r} at another position'
                    assert match[0].default == p.default, f'{ours.__name__}.{meth}: default of {p.name!r} differs ({match[0].default!r} vs {p.default!r})'
