"""CPU: the drop-in boundary — constructor, state_dict layout, C-ABI surface, loud failure without CUDA."""
import ctypes
import json
import os
import re

import pytest
import torch

from golden_util import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name', ['mphoi', 'cad120', 'bimanual'])
def test_state_dict_layout_equals_reference(name, pkg, synth):
    layout = json.load(open(os.path.join(GOLDEN_DIR, 'state_dict_layout.json')))[f'{name}_D32']
    model = pkg.TGGCN(**synth.model_kwargs(synth.SHAPES[name], hidden_size=32, stage=1))
    got = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    assert got == layout


def test_constructor_accepts_yaml_values_and_rejects_the_rest(pkg, synth):
    kw = synth.model_kwargs(synth.MPHOI, hidden_size=64, stage=2)
    m = pkg.TGGCN(**kw)
    assert m.filter_discrete_updates and m.update_segment_threshold == pytest.approx(0.1)
    for bad in (dict(message_type='v1'), dict(attention_style='v1'), dict(message_aggregation='max'),
                dict(object_segment_update_strategy='xyz'), dict(add_time_position=1, time_position_strategy='x'), dict(share_level_mlps=1, bias=False),
                dict(discrete_networks_num_layers=4), dict(message_human_to_objects=False), dict(hidden_size=20)):
        with pytest.raises(NotImplementedError):
            pkg.TGGCN(**{**kw, **bad})
    assert pkg.select_model('2G-GCN') is pkg.TGGCN


def test_level_state_variants_keep_the_reference_layout(pkg, synth):
    """cat_level_states / share_level_mlps (vhoi/models.py:553-570): head shapes and the duplicated state_dict names."""
    D = 32
    cat = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=D, stage=2, cat_level_states=1))
    sd = cat.state_dict()
    assert tuple(sd['human_recognition_mlp.0.weight'].shape) == (10, 4 * D)
    assert tuple(sd['object_prediction_mlp.0.weight'].shape) == (12, 4 * D)
    assert tuple(sd['human_frame_recognition_mlp.0.weight'].shape) == (10, 2 * D)
    share = pkg.TGGCN(**synth.model_kwargs(synth.MPHOI, hidden_size=D, stage=2, share_level_mlps=1))
    sd = share.state_dict()
    assert sd['human_frame_recognition_mlp.0.weight'].data_ptr() == sd['human_recognition_mlp.0.weight'].data_ptr()
    names = [n for n, _ in share.named_parameters()]
    assert 'human_recognition_mlp.0.weight' in names and 'human_frame_recognition_mlp.0.weight' not in names
    both = pkg.TGGCN(**synth.model_kwargs(synth.MPHOI, hidden_size=D, stage=2, share_level_mlps=1, cat_level_states=1))
    assert not both.share_level_mlps and both.cat_level_states          # models.py:565: no sharing with concatenated inputs


def test_update_strategy_variants_keep_the_reference_layout(pkg, synth):
    """object_segment_update_strategy (vhoi/models.py:537-547): 'sah' has no object gate MLP, 'coh' keeps it."""
    ind = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=32, stage=2))
    sah = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=32, stage=2, object_segment_update_strategy='sah'))
    coh = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=32, stage=2, object_segment_update_strategy='conditional_on_human'))
    gate = {'update_object_segment_mlp.0.weight', 'update_object_segment_mlp.0.bias'}
    assert list(coh.state_dict()) == list(ind.state_dict())
    assert [k for k in ind.state_dict() if k not in gate] == list(sah.state_dict())


def test_time_position_variants_keep_the_reference_layout(pkg, synth):
    """add_time_position (vhoi/models.py:259-260, :290, :315, :530, :545): time_position_mlp registered first (embedding only),
    one more D-wide block in the segment cells' inputs ('s') or in the gate MLPs' inputs ('u')."""
    D = 32
    base = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=D, stage=2)).state_dict()
    se = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=D, stage=2, add_time_position=1)).state_dict()
    assert list(se)[:2] == ['time_position_mlp.0.weight', 'time_position_mlp.0.bias'] and list(se)[2:] == list(base)
    assert tuple(se['time_position_mlp.0.weight'].shape) == (D, 1)
    assert se['human_segment_rnn_fcell.weight_ih'].shape[1] == base['human_segment_rnn_fcell.weight_ih'].shape[1] + D
    assert se['object_segment_rnn_bcell.weight_ih'].shape[1] == 7 * D
    assert se['update_object_segment_mlp.0.weight'].shape == base['update_object_segment_mlp.0.weight'].shape
    up = pkg.TGGCN(**synth.model_kwargs(synth.CAD120, hidden_size=D, stage=2, add_time_position=1, time_position_strategy='u',
                                        positional_encoding_style='p')).state_dict()
    assert list(up) == list(base)                                   # periodic: no parameters of its own
    assert up['update_object_segment_mlp.0.weight'].shape[1] == 6 * D and up['update_human_segment_mlp.0.weight'].shape[1] == 4 * D
    assert up['object_segment_rnn_fcell.weight_ih'].shape == base['object_segment_rnn_fcell.weight_ih'].shape


def test_weight_table_covers_the_state_dict(pkg, synth):
    for name in ('mphoi', 'cad120'):
        model = pkg.TGGCN(**synth.model_kwargs(synth.SHAPES[name], hidden_size=32, stage=1))
        keys = set(model.state_dict().keys())
        table = set(pkg.abi.WEIGHT_KEYS)
        assert table <= keys | {k for k in table if k.startswith('object_') or k.startswith('humans_to_human') or k.startswith('time_position_mlp') or k.startswith('geometry_to_human') or k.startswith('segment_length_mlp') or '_segment_mlp.2.' in k or '_segment_mlp.4.' in k}
        used_somewhere = {k for k in keys if k in table}
        # everything not in the table is a parameter the shipped configuration never reads
        dead = keys - used_somewhere
        assert all(('_att_mlp' in k) or k.startswith('geometry_to_object_segment_message_mlp') or
                   k.endswith('num_batches_tracked') for k in dead), sorted(dead)


def test_library_loads_and_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, 'include', 'tggcn_b200.h')).read()
    declared = re.findall(r'TGGCN_API\s+[\w\s\*]+?\b(tggcn_\w+)\s*\(', header)
    assert len(declared) >= 8
    lib = pkg.abi.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.tggcn_abi_version() == pkg.abi.ABI_VERSION == int(re.search(r'#define TGGCN_ABI_VERSION\s+(\d+)', header).group(1))
    # struct mirrors: 17 int32 + 1 float + 8 int32; io = 6 + 4 + 8 + 3 + 3 + 1 pointers
    assert ctypes.sizeof(pkg.abi.Dims) == 32 * 4
    assert ctypes.sizeof(pkg.abi.IO) == 30 * 8
    # the status decoder is host-only: healthy words, a barrier time-out, an fp16-split range violation
    words = (ctypes.c_uint32 * 8)()
    assert lib.tggcn_status_decode(words) == 0
    words[3] = 2
    assert lib.tggcn_status_decode(words) == 2 and b'fp16-split' in lib.tggcn_last_error()
    words[7] = 1
    assert lib.tggcn_status_decode(words) == 3 and b'timed out' in lib.tggcn_last_error()


def test_integration_doc_mirrors_the_structs(pkg):
    """INTEGRATION.md shows the ctypes binding a maintainer would write: its field lists must be the ones of include/tggcn_b200.h
    (= abi.Dims / abi.IO), in order."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = doc[doc.index('class Dims(C.Structure)'):doc.index('KEYS = re.findall')]
    dims_doc = re.findall(r"'(\w+)'", block[:block.index('class IO(C.Structure)')])
    io_doc = re.findall(r"'(\w+)'", block[block.index('class IO(C.Structure)'):])
    assert dims_doc == [n for n, _ in pkg.abi.Dims._fields_]
    assert io_doc == [n for n, _ in pkg.abi.IO._fields_]


def test_workspace_query_needs_no_gpu(pkg):
    d = pkg.abi.Dims(B=8, T=128, H=2, O=4, V=26, D=512, Fh=2152, C_sub=13, C_aff=0, hh=1, filter=1, bn_train=0,
                     human_seg_given=0, object_seg_given=0, inspect=0, persistent=1, gemm_path=0, thr=0.1)
    total = pkg.abi.workspace_bytes(d)
    assert 100e6 < total < 2e9
    off, nbytes = pkg.abi.workspace_view(d, 'GCN_OUT')
    assert off == 0 and nbytes == 8 * 128 * 26 * 128 * 4
    bad = pkg.abi.Dims(B=8, T=128, H=2, O=4, V=26, D=500, Fh=2152, C_sub=13)
    with pytest.raises(pkg.abi.TggcnError):
        pkg.abi.workspace_bytes(bad)


def test_forward_refuses_cpu_tensors(pkg, synth):
    model = pkg.TGGCN(**synth.model_kwargs(synth.MPHOI, hidden_size=32, stage=1))
    batch = synth.make_batch(synth.MPHOI, 2, 5)
    with torch.no_grad(), pytest.raises(pkg.abi.TggcnError):
        model(x_human=batch['x_human'], x_objects=batch['x_objects'], objects_mask=batch['objects_mask'])
