"""Drop-in ``TGGCN`` (the 2G-GCN model class): host-side mirror of ``vhoi/models.py:178-933``.

Same constructor arguments (``conf/models/2G-GCN_stage{1,2}.yaml:4-29`` + ``input_size`` / ``num_classes``,
train.py:28-34), same ``state_dict()`` names and shapes (so reference ``.tar`` checkpoints load both
ways), same keyword-called ``forward`` (vhoi/data_loading.py:1245-1279) and the same output list
(models.py:919-932).  The module only *holds* parameters; all arithmetic of the forward runs in the
hand-written sm_100a kernels behind the C ABI (``include/tggcn_b200.h``).  There is no PyTorch or CPU
fallback: without a CUDA device or without the built library, ``forward`` raises.

Supported configuration family = what the shipped yaml files exercise; every other flag
value raises ``NotImplementedError`` at construction.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Optional

import torch
import torch.nn as nn

from . import abi

_GS = {'gs', 'gumbel-sigmoid'}
_ST = {'st', 'straight-through'}
_ATT = {'att', 'attention'}
_MP = {'mp', 'mean_pooling'}
_V3 = {'v3', 'scaled_dot-product'}
_V2 = {'v2', 'dot-product'}
_NONREL = {'v2', 'non-relational'}
_GENERIC = {'v1', 'generic'}
_IND = {'ind', 'independent'}
_ENC_E = {'e', 'embedding'}
_ENC_P = {'p', 'periodic'}
_SAH = {'sah', 'same_as_human'}
_COH = {'coh', 'conditional_on_human'}


def _mlp(dims, acts):
    """Linear/activation stack with the reference's module indices (pyrutils/torch/models.py:30-36)."""
    layers = []
    for i, a in enumerate(acts):
        layers.append(nn.Linear(dims[i], dims[i + 1], bias=True))
        layers.append({'relu': nn.ReLU, 'sigmoid': nn.Sigmoid}.get(a, nn.Identity)() if a != 'logsoftmax'
                      else nn.LogSoftmax(dim=-1))
    return nn.Sequential(*layers)


class _Holder(nn.Module):
    """Named container used to reproduce the reference's nested parameter paths."""


def _geo_gcn_holder(node_n: int) -> nn.Module:
    """Parameter tree of Geo_gcn (pyrutils/torch/models_gcn.py:16-28): same names, shapes and init."""
    root = _Holder()
    norm = _Holder()
    norm.bn = nn.BatchNorm1d(4 * node_n)
    conv_a, conv_b = _Holder(), _Holder()
    conv_a.cnn = nn.Conv2d(4, 64, kernel_size=1, bias=True)
    conv_b.cnn = nn.Conv2d(64, 64, kernel_size=1, bias=True)
    embed = _Holder()
    embed.cnn = nn.Sequential(norm, conv_a, nn.ReLU(), conv_b, nn.ReLU())
    root.joint_embed = embed
    sim = _Holder()
    sim.s1, sim.s2 = _Holder(), _Holder()
    sim.s1.cnn = nn.Conv2d(64, 128, kernel_size=1, bias=True)
    sim.s2.cnn = nn.Conv2d(64, 128, kernel_size=1, bias=True)
    root.get_s = sim
    root.weight = nn.Parameter(torch.empty(64, 128))
    stdv = 1.0 / math.sqrt(root.weight.size(1))
    with torch.no_grad():
        root.weight.uniform_(-stdv, stdv)
    return root



class _ForwardBackward(torch.autograd.Function):
    """Training-mode bridge: forward = tggcn_forward with dims.save_for_backward, backward = tggcn_backward
    (the hand-written backward kernels).  Inputs after the first two are the parameters of the weight table that
    can receive a gradient, so the unchanged ``loss.backward()`` / ``optimizer.step()`` of
    pyrutils/torch/train_utils.py:150-154 work."""

    @staticmethod
    def forward(ctx, model, launch, *params):
        outputs, state = launch()
        ctx.model, ctx.state = model, state
        ctx.n_out, ctx.n_params = len(outputs), len(params)
        nondiff = [o for o, d in zip(outputs, state['differentiable']) if not d]
        if nondiff:
            ctx.mark_non_differentiable(*nondiff)
        return tuple(outputs)

    @staticmethod
    def backward(ctx, *gouts):
        # _backward binds every parameter's .grad to its slice of the flat buffer itself (no AccumulateGrad copy of 182 MB, and
        # the data-parallel reducer can start on a bucket while later stages still run); autograd gets no tensors to accumulate
        ctx.model._backward(ctx.state, gouts)
        return (None, None) + (None,) * ctx.n_params


class OutputList(list):
    """The output list of a forward (vhoi/models.py:919-932).  In inference ``flat`` is the one contiguous fp32 buffer all output
    tensors are views of (gates first, then the heads; 256-byte aligned pieces), so a caller that wants everything on the host can do
    it with a single device->host copy instead of one per tensor; ``None`` under autograd."""
    flat = None

    def __reduce__(self):          # pickles (torch.save of a prediction list) as the plain list the reference returns
        return (list, (list(self),))


class TGGCN(nn.Module):
    """B200-native 2G-GCN.  See module docstring; argument meaning as in vhoi/models.py:191-233."""
    _STATUS_SLOTS = 64

    def __init__(self, input_size: tuple, num_classes: tuple, hidden_size: int = 128,
                 discrete_networks_num_layers: int = 1, discrete_optimization_strategy: str = 'gumbel-sigmoid',
                 filter_discrete_updates: bool = False, gcn_node: int = 26,
                 message_humans_to_human: bool = True, message_human_to_objects: bool = True,
                 message_objects_to_human: bool = True, message_objects_to_object: bool = True,
                 message_geometry_to_objects: bool = True, message_geometry_to_human: bool = False,
                 message_segment: bool = False, message_type: str = 'relational', message_granularity: str = 'specific',
                 message_aggregation: str = 'attention', attention_style: str = 'concat',
                 object_segment_update_strategy: str = 'independent', update_segment_threshold: float = 0.5,
                 add_segment_length: bool = False, add_time_position: bool = False, time_position_strategy: str = 's',
                 positional_encoding_style: str = 'embedding', cat_level_states: bool = False,
                 share_level_mlps: bool = False, bias: bool = True):
        super().__init__()
        unsupported = []
        if discrete_networks_num_layers not in (1, 2, 3): unsupported.append('discrete_networks_num_layers not in {1, 2, 3}')
        if discrete_optimization_strategy not in _GS | _ST: unsupported.append('unknown discrete_optimization_strategy')
        if not (message_human_to_objects and message_objects_to_human and message_objects_to_object
                and message_geometry_to_objects): unsupported.append('a human/object/geometry message switched off')
        if not message_segment: unsupported.append('message_segment off')
        if message_type not in _NONREL: unsupported.append("message_type != 'v2'")
        if message_granularity not in _GENERIC: unsupported.append("message_granularity != 'v1'")
        if message_aggregation not in _ATT | _MP: unsupported.append("message_aggregation not in {'att', 'mp'}")
        if attention_style not in _V3 | _V2: unsupported.append("attention_style not in {'v2', 'v3'}")
        if object_segment_update_strategy not in _IND | _SAH | _COH: unsupported.append('unknown object_segment_update_strategy')
        if add_time_position and time_position_strategy not in ('s', 'u'): unsupported.append("time_position_strategy not in {'s', 'u'}")
        if (add_time_position or add_segment_length) and positional_encoding_style not in _ENC_E | _ENC_P: unsupported.append('unknown positional_encoding_style')
        if not bias: unsupported.append('bias=False')
        if hidden_size % 16 != 0: unsupported.append('hidden_size not a multiple of 16')
        if unsupported:
            raise NotImplementedError('2G-GCN B200 path supports the shipped configuration family only; got: '
                                      + '; '.join(unsupported))
        human_input_size, object_input_size = input_size
        if human_input_size != 2048 + 4 * gcn_node or object_input_size != 2048:
            raise NotImplementedError(f'input_size {tuple(input_size)} does not match 2048+4*gcn_node / 2048')
        n_sub, n_aff = num_classes
        D = hidden_size
        # attributes the reference also keeps (vhoi/models.py:237-257)
        self.discrete_optimization_strategy = discrete_optimization_strategy
        self.filter_discrete_updates = bool(filter_discrete_updates)
        self.gcn_node = gcn_node
        self.message_humans_to_human = bool(message_humans_to_human)
        self.message_human_to_objects = True
        self.message_objects_to_human = True
        self.message_objects_to_object = True
        self.message_geometry_to_objects = True
        self.message_geometry_to_human = bool(message_geometry_to_human)
        self.message_segment = True
        self.message_type, self.message_granularity = message_type, message_granularity
        self.message_aggregation, self.attention_style = message_aggregation, attention_style
        self.object_segment_update_strategy = object_segment_update_strategy
        self.update_segment_threshold = float(update_segment_threshold)
        self.add_segment_length, self.add_time_position = bool(add_segment_length), bool(add_time_position)
        self.time_position_strategy, self.positional_encoding_style = time_position_strategy, positional_encoding_style
        self.cat_level_states = bool(cat_level_states)
        self.share_level_mlps = bool(share_level_mlps) and not self.cat_level_states     # models.py:565: sharing needs equal input sizes
        self.hidden_size, self.num_classes = D, (n_sub, n_aff)
        hh = self.message_humans_to_human
        # time-position features (models.py:259-260, :290, :315, :530, :545): one more D-wide block in the segment-level inputs
        # (strategy 's') or in the gate inputs ('u')
        ts = int(self.add_time_position and time_position_strategy == 's')
        tu = int(self.add_time_position and time_position_strategy == 'u')
        self._time_periodic = positional_encoding_style in _ENC_P
        # ---- parameter holders, registered in the reference's order (vhoi/models.py:264-580) ----------
        if self.add_time_position and not self._time_periodic:
            self.time_position_mlp = _mlp([1, D], ['relu'])
        if self.add_segment_length and not self._time_periodic:       # models.py:261-262
            self.segment_length_mlp = _mlp([1, D], ['relu'])
        ts += int(self.add_segment_length)                             # models.py:292-293, :317-318: one more D-wide input block
        self.geometry_embedding_gcn = _geo_gcn_holder(gcn_node)
        self.geometry_embedding_mlp = _mlp([gcn_node * 128, 2048, D], ['relu', 'relu'])
        self.geometry_bd_rnn = nn.GRU(D, D, num_layers=1, bias=True, batch_first=True, bidirectional=True)
        self.geometry_bd_embedding_mlp = _mlp([2 * D, D], ['relu'])
        self.human_embedding_mlp = _mlp([2048, D], ['relu'])
        self.human_bd_rnn = nn.GRU(D, D, num_layers=1, bias=True, batch_first=True, bidirectional=True)
        self.human_bd_embedding_mlp = _mlp([2 * D, D], ['relu'])
        gh = int(self.message_geometry_to_human)
        h_in = D * (1 + (2 if hh else 0) + 2 + gh + ts)
        self.human_segment_rnn_fcell = nn.GRUCell(h_in, D, bias=True)
        self.human_segment_rnn_bcell = nn.GRUCell(h_in, D, bias=True)
        self.object_embedding_mlp = _mlp([object_input_size, D], ['relu'])
        self.object_bd_rnn = nn.GRU(D, D, num_layers=1, bias=True, batch_first=True, bidirectional=True)
        self.object_bd_embedding_mlp = _mlp([2 * D, D], ['relu'])
        self.object_segment_rnn_fcell = nn.GRUCell((6 + ts) * D, D, bias=True)
        self.object_segment_rnn_bcell = nn.GRUCell((6 + ts) * D, D, bias=True)
        kinds = (['humans_to_human'] if hh else []) + ['human_to_object', 'objects_to_human', 'objects_to_object']
        att_names = {'humans_to_human': 'humans_to_human', 'human_to_object': 'humans_to_object',
                     'objects_to_human': 'objects_to_human', 'objects_to_object': 'objects_to_object'}
        for kind in kinds:
            setattr(self, f'{kind}_message_mlp', _mlp([2 * D, D], ['relu']))
            setattr(self, f'{kind}_segment_message_mlp', _mlp([D, D], ['relu']))
            # present in the reference's state_dict (attention aggregation only) but unused under attention_style 'v3'
            if message_aggregation in _ATT:
                setattr(self, f'{att_names[kind]}_message_att_mlp', _mlp([4 * D, 1], ['relu']))
                setattr(self, f'{att_names[kind]}_segment_message_att_mlp', _mlp([2 * D, 1], ['relu']))
        if gh:                                                     # models.py:456-488; the segment-level MLP and the att MLPs are dead
            self.geometry_to_human_message_mlp = _mlp([2 * D, D], ['relu'])
            self.geometry_to_human_segment_message_mlp = _mlp([D, D], ['relu'])
            if message_aggregation in _ATT:
                self.geometry_to_human_message_att_mlp = _mlp([4 * D, 1], ['relu'])
                self.geometry_to_human_segment_message_att_mlp = _mlp([2 * D, 1], ['relu'])
        self.geometry_to_object_message_mlp = _mlp([2 * D, D], ['relu'])
        self.geometry_to_object_segment_message_mlp = _mlp([D, D], ['relu'])          # dead in the reference too
        if message_aggregation in _ATT:
            self.geometry_to_object_message_att_mlp = _mlp([4 * D, 1], ['relu'])
            self.geometry_to_object_segment_message_att_mlp = _mlp([2 * D, 1], ['relu'])
        self.discrete_networks_num_layers = int(discrete_networks_num_layers)
        gate_hidden = [D] * (self.discrete_networks_num_layers - 1)          # models.py:532-535
        gate_act = ['relu'] * len(gate_hidden) + ['sigmoid']
        self.update_human_segment_mlp = _mlp([D * (2 + (1 if hh else 0) + 1 + gh + tu)] + gate_hidden + [1], gate_act)
        if object_segment_update_strategy not in _SAH:            # models.py:537: no object gate MLP under 'sah'
            self.update_object_segment_mlp = _mlp([(5 + tu) * D] + gate_hidden + [1], gate_act)
        label_in = (4 if self.cat_level_states else 2) * D        # models.py:553-555
        self.human_recognition_mlp = _mlp([label_in, n_sub], ['logsoftmax'])
        self.human_prediction_mlp = _mlp([label_in, n_sub], ['logsoftmax'])
        if n_aff is not None:
            self.object_recognition_mlp = _mlp([label_in, n_aff], ['logsoftmax'])
            self.object_prediction_mlp = _mlp([label_in, n_aff], ['logsoftmax'])
        if self.share_level_mlps:                                  # models.py:565-570: the SAME modules under a second name
            self.human_frame_recognition_mlp = self.human_recognition_mlp
            self.human_frame_prediction_mlp = self.human_prediction_mlp
            if n_aff is not None:
                self.object_frame_recognition_mlp = self.object_recognition_mlp
                self.object_frame_prediction_mlp = self.object_prediction_mlp
        else:
            self.human_frame_recognition_mlp = _mlp([2 * D, n_sub], ['logsoftmax'])
            self.human_frame_prediction_mlp = _mlp([2 * D, n_sub], ['logsoftmax'])
            if n_aff is not None:
                self.object_frame_recognition_mlp = _mlp([2 * D, n_aff], ['logsoftmax'])
                self.object_frame_prediction_mlp = _mlp([2 * D, n_aff], ['logsoftmax'])
        # ---- runtime state (not part of state_dict) ------------------------------------------------------
        self._ptr_cache = None
        self._dir = None                    # cached (owner module, attribute, is_parameter, name, tensor) of every state_dict entry
        self._dir_version = 0
        self._trainable_cache = {}
        self._ws = {}
        self._time_freq = None
        self._noise_ring = {}               # (n_calls, B) -> pinned buffers of the default Gumbel draws
        self._noise_override: Optional[torch.Tensor] = None
        self._generation = 0
        self.flat_grad = None
        self.persistent_kernels = True      # False: one launch per recurrent step (debug aid)
        self.gemm_path = 2                  # 0: fp32 SIMT projections; 1: tcgen05 3xTF32; 2: tcgen05 where K % 32 == 0
        self.recurrent_mode = int(os.environ.get('TGGCN_RECURRENT_MODE', '0'))   # dims.recurrent_mode (0 = by rows per step)
        self.no_fp16_split = False          # set after a range violation: 3xTF32 streaming recurrent kernels from then on
        self.precision = 0                  # dims.precision: 0 = fp32-class split products, 1 = bf16 operands (set_precision)
        self._pending_status = []           # (event, ring slot, what) of calls whose status words have not been looked at
        self._status_ring = None            # pinned (slots, 8) int32: landing zones of the asynchronous status copies
        self._status_next = 0
        self._bucket_events = {}            # device -> cudaEvents recorded by tggcn_backward_ex at the gradient-bucket boundaries
        self._flat_pad_index = None
        self.grad_buckets = []              # [(start, end, event)] slices of flat_grad in completion order (last backward)
        self.grad_ready_callback = None     # called with the model once a backward has been queued (data-parallel reducer)

    # ------------------------------------------------------------------------------------------------
    def set_gumbel_noise(self, noise: Optional[torch.Tensor]):
        """Inject the Gumbel(0,1) draws of the next forward calls: (n_calls, B, 2) in the reference's call
        order (t-major; sampled humans, then sampled objects).  ``None`` restores the default, which draws
        from the global CPU generator exactly like pyrutils/torch/distributions.py:16 does."""
        self._noise_override = noise

    def set_precision(self, precision: str):
        """'fp32' (default): every matrix product is a 3-term split MMA with fp32-class accuracy — the configuration all parity
        tests run.  'bf16': projections and the large-batch recurrent kernels round their operands to bf16 and accumulate in
        fp32 (BASELINE.json configs[2], "training step bf16"); gates, softmaxes, losses and the optimiser state stay fp32."""
        if precision not in ('fp32', 'bf16'):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = 1 if precision == 'bf16' else 0
        return self

    # -- status words of the persistent kernels -------------------------------------------------------------------------
    def _queue_status(self, io, dev, what):
        """Give the call a pinned landing zone for its status words; they are looked at by a LATER call (or by
        check_persistent_kernels), so the check costs no synchronisation on the hot path."""
        if self._status_ring is None:           # one pinned allocation for the life of the model (cudaHostAlloc is slow and may synchronise)
            self._status_ring = torch.zeros(self._STATUS_SLOTS, 8, dtype=torch.int32).pin_memory()
        slot = self._status_next % self._STATUS_SLOTS
        self._status_next += 1
        if any(s == slot for _, s, _ in self._pending_status):      # the ring wrapped onto a call nobody has looked at: look now
            self._poll_status(wait=True)
        words = self._status_ring[slot]
        io.status_host = words.data_ptr()
        return (torch.cuda.Event(), slot, what)

    def _poll_status(self, wait: bool = False):
        """Test the status words of earlier calls whose copy has landed (all of them when ``wait``).  A grid-barrier time-out
        or an fp16-split range violation raises — results of that call were wrong; after a range violation the model switches
        itself to the 3xTF32 streaming kernels, so a caller that catches the error can simply repeat the step."""
        still = []
        err = None
        for ev, slot, what in self._pending_status:
            if wait:
                ev.synchronize()
            if not ev.query():
                still.append((ev, slot, what))
                continue
            rc = abi.lib().tggcn_status_decode(C.c_void_p(self._status_ring[slot].data_ptr()))
            if rc != 0 and err is None:
                msg = abi.lib().tggcn_last_error().decode(errors='replace')
                if rc & 2:
                    self.no_fp16_split = True
                    msg += ' — this model now runs the 3xTF32 streaming recurrent kernels (no_fp16_split); repeat the step'
                err = abi.TggcnError(f'{what}: {msg}')
        self._pending_status = still
        if err is not None:
            raise err

    def _apply(self, fn, *args, **kwargs):
        self._ptr_cache = None
        self._ws = {}
        return super()._apply(fn, *args, **kwargs)

    def _tensor_directory(self):
        """[(state_dict name, tensor, is_parameter)] in named_parameters() + named_buffers() order, cached: walking the module tree
        costs ~0.7 ms per call and a training step used to do it four times.  The cache is validated by identity — every entry's
        owner module must still hold that very tensor object under that attribute (~130 dict look-ups) — so replacing a parameter,
        or the copies _apply() makes, rebuild it."""
        d = self._dir
        if d is not None and all((m._parameters if isp else m._buffers).get(a) is t for m, a, isp, _, t in d):
            return d
        d = []
        for isp, items in ((True, self.named_parameters()), (False, self.named_buffers())):
            for name, t in items:
                owner, _, attr = name.rpartition('.')
                d.append((self.get_submodule(owner) if owner else self, attr, isp, name, t))
        self._dir = d
        self._dir_version += 1
        self._trainable_cache = {}
        self._ptr_cache = None
        return d

    def _weight_pointers(self, device):
        sd_items = [(name, t) for _, _, _, name, t in self._tensor_directory()]
        # every storage address takes part (~20 us): a parameter whose .data was re-pointed (optim.FlatAdam, user code) must not be
        # read through a stale pointer
        probe = (hash(tuple(t.data_ptr() for _, t in sd_items)), len(sd_items), self._dir_version)
        if self._ptr_cache is not None and self._ptr_cache[0] == probe:
            return self._ptr_cache[1]
        arr = (C.c_void_p * abi.N_WEIGHTS)()
        for name, t in sd_items:
            idx = abi.WEIGHT_INDEX.get(name)
            if idx is None:
                continue
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                raise abi.TggcnError(f'parameter {name} must be a contiguous fp32 tensor on {device}')
            arr[idx] = t.data_ptr()
        self._alias_shared_heads(arr)
        self._ptr_cache = (probe, arr)
        return arr

    def _alias_shared_heads(self, table):
        """share_level_mlps: the frame-level heads ARE the segment-level heads (vhoi/models.py:565-570); named_parameters()
        lists them once, under the segment-level name, so the frame-level slots of a pointer table point at the same memory."""
        if not self.share_level_mlps:
            return
        for ent in ('human', 'object'):
            for kind in ('recognition', 'prediction'):
                for leaf in ('0.weight', '0.bias'):
                    src = abi.WEIGHT_INDEX.get(f'{ent}_{kind}_mlp.{leaf}')
                    dst = abi.WEIGHT_INDEX.get(f'{ent}_frame_{kind}_mlp.{leaf}')
                    if src is not None and dst is not None and table[src]:
                        table[dst] = table[src]

    def _workspace(self, dims: abi.Dims, device, backward: bool = False):
        key = (dims.B, dims.T, dims.H, dims.O, dims.save_for_backward, backward, device)
        ws = self._ws.get(key)
        if ws is None:
            if backward:
                nbytes = int(abi.lib().tggcn_backward_workspace_bytes(C.byref(dims)))
                if nbytes == 0:
                    raise abi.TggcnError('tggcn_backward_workspace_bytes: ' + abi.lib().tggcn_last_error().decode(errors='replace'))
            else:
                nbytes = abi.workspace_bytes(dims)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            if len(self._ws) > 8:
                self._ws.clear()
            self._ws[key] = ws
        return ws

    @staticmethod
    def draw_gumbel_noise(n_calls: int, batch: int) -> torch.Tensor:
        """All Gumbel(0,1) draws of one forward in one CPU call; bit-identical to n_calls successive
        ``Gumbel(0,1).sample((batch, 2))`` calls on the global CPU generator (tests/test_noise.py)."""
        fi = torch.finfo(torch.float32)
        u = torch.rand(n_calls, batch, 2)
        u = u * ((1.0 - fi.eps) - fi.tiny) + fi.tiny
        return -torch.log(-torch.log(u))

    def _draw_noise_pinned(self, n_calls: int, batch: int):
        """The default noise path of a forward: the same draws as ``draw_gumbel_noise`` (bit for bit, same generator stream), made
        in place in a small ring of PINNED host buffers which the gate kernel reads DIRECTLY (pinned memory is device-accessible
        under unified addressing; 2 floats per sampled gate).  No host->device copy is queued at all: a copy from pageable memory
        makes CUDA synchronise the stream first (the host would stall behind the previous forward on every call), and a pinned
        asynchronous copy queues on the copy engine behind whatever batch prefetch is in flight (measured: +0.8 ms per forward).
        Returns (pinned tensor, ring, slot); the caller records ring['events'][slot] once the forward has been queued."""
        key = (n_calls, batch)
        ring = self._noise_ring.get(key)
        if ring is None:
            if len(self._noise_ring) > 4:
                self._noise_ring.clear()
            ring = self._noise_ring[key] = dict(bufs=[torch.empty(n_calls, batch, 2).pin_memory() for _ in range(3)],
                                                events=[None] * 3, next=0)
        i = ring['next']
        ring['next'] = (i + 1) % 3
        if ring['events'][i] is not None:
            ring['events'][i].synchronize()         # the forward that last read this buffer (three calls ago) has finished
        fi = torch.finfo(torch.float32)
        u = ring['bufs'][i]
        torch.rand(n_calls, batch, 2, out=u)
        u.mul_((1.0 - fi.eps) - fi.tiny).add_(fi.tiny)
        u.log_().neg_().log_().neg_()
        return u, ring, i

    # ------------------------------------------------------------------------------------------------
    def forward(self, x_human, x_objects, objects_mask, human_segmentation=None, objects_segmentation=None,
                human_human_distances=None, human_object_distances=None, object_object_distances=None,
                steps_per_example=None, inspect_model=False):
        """Same contract as vhoi/models.py:584-623.  Returns the list of models.py:919-926 (6 tensors, or 12
        when affordance classes exist), plus the attention stacks when ``inspect_model``."""
        return self._run(x_human, x_objects, objects_mask, human_segmentation, objects_segmentation,
                         human_human_distances, human_object_distances, object_object_distances, inspect_model, None,
                         steps_per_example)

    def forward_profile(self, x_human, x_objects, objects_mask, human_segmentation=None, objects_segmentation=None,
                        inspect_model=False, steps_per_example=None):
        """forward() with CUDA events around every stage; returns (outputs, {stage name: milliseconds})."""
        ms = (C.c_float * len(abi.STAGE_NAMES))()
        out = self._run(x_human, x_objects, objects_mask, human_segmentation, objects_segmentation, None, None, None,
                        inspect_model, ms, steps_per_example)
        return out, dict(zip(abi.STAGE_NAMES, list(ms)))

    def _run(self, x_human, x_objects, objects_mask, human_segmentation, objects_segmentation, hh_d, ho_d, oo_d,
             inspect_model, stage_ms, steps_per_example=None):
        if not x_human.is_cuda:
            raise abi.TggcnError('2G-GCN B200 path runs on a CUDA device only (no CPU fallback); got a CPU tensor')
        self._poll_status()                 # status words of earlier calls that have landed by now
        dev = x_human.device
        B, T, H, Fh = x_human.shape
        O = x_objects.size(2)
        n_sub, n_aff = self.num_classes
        f32 = dict(dtype=torch.float32, device=dev)
        with_grad = torch.is_grad_enabled() and any(t.requires_grad for _, _, isp, _, t in self._tensor_directory() if isp)
        if inspect_model and self.message_aggregation in _MP:
            raise NotImplementedError('inspect_model has no attention weights to return under mean-pooling aggregation')
        if with_grad and self.discrete_optimization_strategy in _ST and (human_segmentation is None or objects_segmentation is None):
            raise NotImplementedError("discrete_optimization_strategy 'st' is inference-only: the reference's StraightThroughEstimator.backward "
                                      'returns one gradient for two inputs and autograd rejects it (pyrutils/torch/distributions.py:39-53)')
        if with_grad and (inspect_model or stage_ms is not None):
            raise NotImplementedError('inspect_model / stage profiling are inference-only: call under torch.no_grad()')

        def prep(t):
            return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()
        x_human, x_objects, objects_mask = prep(x_human), prep(x_objects), prep(objects_mask)
        hseg = prep(human_segmentation) if human_segmentation is not None else None
        oseg = prep(objects_segmentation) if objects_segmentation is not None else None

        dims = abi.Dims(B=B, T=T, H=H, O=O, V=self.gcn_node, D=self.hidden_size, Fh=Fh, C_sub=n_sub,
                        C_aff=0 if n_aff is None else n_aff, hh=int(self.message_humans_to_human),
                        filter=int(self.filter_discrete_updates), bn_train=int(self.training),
                        human_seg_given=int(hseg is not None), object_seg_given=int(oseg is not None),
                        inspect=int(bool(inspect_model)), persistent=int(self.persistent_kernels),
                        gemm_path=int(self.gemm_path), thr=self.update_segment_threshold,
                        save_for_backward=int(with_grad), cat_level_states=int(self.cat_level_states),
                        mean_pool=int(self.message_aggregation in _MP), recurrent_mode=int(self.recurrent_mode),
                        no_fp16_split=int(self.no_fp16_split), precision=int(self.precision),
                        att_noscale=int(self.attention_style in _V2))
        dims.straight_through = int(self.discrete_optimization_strategy in _ST)
        dims.geo_to_human = int(self.message_geometry_to_human)
        dims.gate_layers = self.discrete_networks_num_layers
        # misc.make_attention_distance_based (data_loading.py:1264-1276): meaningful under attention aggregation only (models.py:1033-1046)
        dists = [None, None, None]
        if self.message_aggregation in _ATT:
            want = ((B, T, H, H), (B, T, H, O), (B, T, O, O))
            for i, (t_, shp) in enumerate(zip((hh_d, ho_d, oo_d), want)):
                if t_ is not None:
                    if tuple(t_.shape) != shp:
                        raise ValueError(f'distance tensor {i} must have shape {shp}, got {tuple(t_.shape)}')
                    dists[i] = t_.to(device=dev, dtype=torch.float32).contiguous()
        steps = freq = None
        if self.add_time_position or self.add_segment_length:
            dims.time_position = (1 if self.time_position_strategy == 's' else 2) if self.add_time_position else 0
            dims.segment_length = int(self.add_segment_length)
            dims.time_periodic = int(self._time_periodic)
            if steps_per_example is None:
                raise ValueError('add_time_position / add_segment_length need steps_per_example (vhoi/data_loading.py:1277)')
            steps = steps_per_example.to(device=dev, dtype=torch.float32).contiguous()
            if self._time_periodic:             # the period table of make_periodic_embedding (models.py:1788-1790), a constant of D
                if self._time_freq is None or self._time_freq.device != dev:
                    self._time_freq = (torch.tensor([1e4]) ** torch.linspace(0, 1, self.hidden_size // 2)).to(dev)
                freq = self._time_freq
        # object_segment_update_strategy (models.py:741-742, :1523-1532): 'sah' / 'coh' act with exactly one human; with more the
        # reference falls back to 'ind' ('sah' then has no object gate MLP to fall back on and fails there too)
        strat = 1 if self.object_segment_update_strategy in _SAH else 2 if self.object_segment_update_strategy in _COH else 0
        if strat == 1 and H != 1 and oseg is None:
            raise ValueError("object_segment_update_strategy 'sah' needs exactly one human (the reference has no object gate MLP)")
        if strat and H == 1 and oseg is None:
            if hseg is not None:
                raise NotImplementedError("object_segment_update_strategy 'sah'/'coh' with an imposed human_segmentation: the "
                                          "reference writes the last step's 1.0 into the caller's tensor (models.py:744-745)")
            if strat == 1 or not self.filter_discrete_updates:       # under the filter 'coh' equals 'ind' (models.py:751-753)
                dims.update_strategy = strat
        objects_sampled = oseg is None and dims.update_strategy != 1
        n_sampled = 0 if dims.straight_through else (0 if hseg is not None else H) + (O if objects_sampled else 0)
        noise = noise_ring = None
        if n_sampled:
            noise = self._noise_override
            if noise is None:
                noise, noise_ring, noise_slot = self._draw_noise_pinned(T * n_sampled, B)
            else:
                if tuple(noise.shape) != (T * n_sampled, B, 2):
                    raise ValueError(f'gumbel noise must have shape {(T * n_sampled, B, 2)}, got {tuple(noise.shape)}')
                noise = noise.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()

        def launch():
            shapes = [(B, T, H), (B, T, H), (B, T, O), (B, T, O)] + [(B, n_sub, T, H)] * 4 + ([(B, n_aff, T, O)] * 4 if n_aff is not None else [])
            flat = None
            if with_grad:                           # autograd outputs: separate allocations
                bufs = [torch.empty(*s, **f32) for s in shapes]
            else:                                   # inference: one allocation, so that a caller can fetch everything with ONE copy
                sizes = [(math.prod(s) + 63) // 64 * 64 for s in shapes]         # every output starts on a 256-byte boundary
                flat = torch.empty(sum(sizes), **f32)
                bufs, off = [], 0
                for s, n in zip(shapes, sizes):
                    bufs.append(flat[off:off + math.prod(s)].view(s))
                    off += n
            y_hs, y_hss, y_os, y_oss = bufs[:4]
            out_h, out_o = bufs[4:8], bufs[8:12]
            att = [torch.zeros(B, H, T, O, **f32) for _ in range(3)] if inspect_model else []

            io = abi.IO()
            io.x_human, io.x_objects, io.objects_mask = x_human.data_ptr(), x_objects.data_ptr(), objects_mask.data_ptr()
            io.human_seg = hseg.data_ptr() if hseg is not None else None
            io.object_seg = oseg.data_ptr() if oseg is not None else None
            io.noise = noise.data_ptr() if noise is not None else None
            io.steps_per_example = steps.data_ptr() if steps is not None else None
            io.dist_hh, io.dist_ho, io.dist_oo = (t_.data_ptr() if t_ is not None else None for t_ in dists)
            io.time_freq = freq.data_ptr() if freq is not None else None
            io.y_hs, io.y_hss, io.y_os, io.y_oss = y_hs.data_ptr(), y_hss.data_ptr(), y_os.data_ptr(), y_oss.data_ptr()
            for i in range(4):
                io.out_h[i] = out_h[i].data_ptr()
                io.out_o[i] = out_o[i].data_ptr() if out_o else None
            if inspect_model:
                io.att_frame, io.att_seg_f, io.att_seg_b = (a.data_ptr() for a in att)
            bn = self.geometry_embedding_gcn.joint_embed.cnn[0].bn
            io.bn_running_mean, io.bn_running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            io.bn_num_batches = bn.num_batches_tracked.data_ptr()

            ws = self._workspace(dims, dev)
            weights = self._weight_pointers(dev)
            pending = self._queue_status(io, dev, 'forward')
            with torch.cuda.device(dev):
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                if stage_ms is None:
                    rc = abi.lib().tggcn_forward(C.byref(dims), weights, abi.N_WEIGHTS, C.byref(io), ws.data_ptr(),
                                                 ws.numel(), stream)
                else:
                    rc = abi.lib().tggcn_forward_profile(C.byref(dims), weights, abi.N_WEIGHTS, C.byref(io), ws.data_ptr(),
                                                         ws.numel(), stream, stage_ms)
            abi.check(rc, 'tggcn_forward')
            if noise_ring is not None:              # the pinned noise buffer is free again once this forward has run
                noise_ring['events'][noise_slot] = torch.cuda.Event()
                noise_ring['events'][noise_slot].record(torch.cuda.current_stream(dev))
            pending[0].record(torch.cuda.current_stream(dev))
            self._pending_status.append(pending)
            # keep inputs alive until the queued work ran
            keep = (x_human, x_objects, objects_mask, hseg, oseg, noise, steps, freq, dists)
            self._last = (dims, ws, keep)
            if n_aff is None:
                output = [y_hs, y_hss] + out_h
                diff = [hseg is None, hseg is None] + [True] * 4
            else:
                output = [y_hs, y_os, y_hss, y_oss] + out_h[:2] + out_o[:2] + out_h[2:] + out_o[2:]
                diff = [hseg is None, oseg is None, hseg is None, oseg is None] + [True] * 8
            output = OutputList(output)
            output.flat = flat
            self._generation += 1
            state = dict(dims=dims, io=io, ws=ws, keep=keep, gates=(y_hs, y_hss, y_os, y_oss), differentiable=diff,
                         generation=self._generation)
            return output, (att if inspect_model else state)

        if not with_grad:
            output, att = launch()
            return (output, att) if inspect_model else output
        names, params = self._trainable_table(dims)
        self._grad_names = names
        return list(_ForwardBackward.apply(self, launch, *params))

    def _trainable_table(self, dims):
        """Parameters of the weight table that are on the gradient path of this call (the others keep grad=None,
        like the 22-24 dead tensors of the reference)."""
        directory = self._tensor_directory()
        key = (dims.human_seg_given, dims.object_seg_given, dims.time_position)
        hit = self._trainable_cache.get(key)
        if hit is not None:
            return hit
        names, params = [], []
        for _, _, isp, name, prm in directory:
            if not isp:
                continue
            if name not in abi.WEIGHT_INDEX:
                continue                       # *_att_mlp, geometry_to_object_segment_message_mlp: never used
            if name.startswith('update_human_segment_mlp') and dims.human_seg_given:
                continue
            if name.startswith('update_object_segment_mlp') and dims.object_seg_given:
                continue
            if name.startswith('time_position_mlp') and dims.time_position == 2 and dims.human_seg_given and dims.object_seg_given:
                continue                       # strategy 'u' feeds the gate MLPs only
            names.append(name)
            params.append(prm)
        self._trainable_cache[key] = (names, params)
        return names, params

    def _backward(self, state, gouts):
        """tggcn_backward on the workspace of the matching forward.  Returns one gradient per tensor of
        ``_trainable_table`` (views into one flat buffer, ``self.flat_grad``, so that a data-parallel
        driver can all-reduce it in a single call)."""
        if state['generation'] != self._generation:
            raise abi.TggcnError('backward called after a later training-mode forward reused the workspace; '
                                 'run forward -> backward one at a time')
        dims, io, ws = state['dims'], state['io'], state['ws']
        dev = ws.device
        names, params = self._trainable_table(dims)
        # flat layout: bucket by bucket in the order the backward completes them (tggcn_backward_bucket), so a data-parallel
        # reducer can all-reduce bucket k as soon as its event has been recorded
        bucket_of = [abi.lib().tggcn_backward_bucket(abi.WEIGHT_INDEX[n]) for n in names]
        order = sorted(range(len(names)), key=lambda i: (bucket_of[i], i))
        offs, total = [0] * len(names), 0
        bucket_ranges = []
        for k in range(abi.BWD_BUCKETS):
            start = total
            for i in order:
                if bucket_of[i] == k:
                    offs[i] = total
                    total += (params[i].numel() + 63) // 64 * 64     # every gradient (and, under optim.FlatAdam, every parameter) on a
                                                                     # 256-byte boundary: the kernels read the weights in place
            total = (total + 63) // 64 * 64                          # buckets start on 256-byte boundaries
            bucket_ranges.append((start, total))
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        sizes = [p.numel() for p in params]
        key = (total, tuple(offs), str(dev))
        if self._flat_pad_index is None or self._flat_pad_index[0] != key:
            used = sorted(zip(offs, sizes))
            gaps, pos = [], 0
            for o, n in used:
                gaps.extend(range(pos, o))
                pos = o + n
            gaps.extend(range(pos, total))
            self._flat_pad_index = (key, torch.tensor(gaps, dtype=torch.int64, device=dev))
        if self._flat_pad_index[1].numel():
            flat.index_fill_(0, self._flat_pad_index[1], 0.0)        # alignment gaps take part in the all-reduce: keep them finite
        grads = [flat[o:o + n].view(p.shape) for o, n, p in zip(offs, sizes, params)]
        garr = (C.c_void_p * abi.N_WEIGHTS)()
        for name, g in zip(names, grads):
            garr[abi.WEIGHT_INDEX[name]] = g.data_ptr()
        self._alias_shared_heads(garr)           # shared heads: both uses accumulate into the one gradient
        n_aff = self.num_classes[1]

        def gp(t):
            if t is None:
                return None, None
            t = t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()
            return t, t.data_ptr()
        go = abi.GradOutputs()
        keep = []
        if n_aff is None:
            order = {'d_y_hs': 0, 'd_y_hss': 1}
            heads_h, heads_o = [2, 3, 4, 5], []
        else:
            order = {'d_y_hs': 0, 'd_y_os': 1, 'd_y_hss': 2, 'd_y_oss': 3}
            heads_h, heads_o = [4, 5, 8, 9], [6, 7, 10, 11]
        for field, i in order.items():
            if state['differentiable'][i]:
                t, ptr = gp(gouts[i])
                keep.append(t)
                setattr(go, field, ptr)
        for j, i in enumerate(heads_h):
            t, ptr = gp(gouts[i])
            keep.append(t)
            go.d_out_h[j] = ptr
        for j, i in enumerate(heads_o):
            t, ptr = gp(gouts[i])
            keep.append(t)
            go.d_out_o[j] = ptr
        bws = self._workspace(dims, dev, backward=True)
        weights = self._weight_pointers(dev)
        pending = self._queue_status(io, dev, 'backward')
        hooks = abi.BwdHooks()
        events = self._bucket_events.get(dev)
        if events is None:
            events = [torch.cuda.Event() for _ in range(abi.BWD_BUCKETS)]
            for ev in events:
                ev.record(torch.cuda.current_stream(dev))            # torch creates the cudaEvent_t lazily, at the first record
            self._bucket_events[dev] = events
        for k, ev in enumerate(events):
            hooks.bucket_done[k] = ev.cuda_event
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = abi.lib().tggcn_backward_ex(C.byref(dims), weights, garr, abi.N_WEIGHTS, C.byref(io), C.byref(go), ws.data_ptr(),
                                             ws.numel(), bws.data_ptr(), bws.numel(), stream, C.byref(hooks))
        abi.check(rc, 'tggcn_backward')
        pending[0].record(torch.cuda.current_stream(dev))
        self._pending_status.append(pending)
        self._last_bwd = (keep, bws)
        self.flat_grad, self._flat_views = flat, (params, grads)
        self._flat_layout = (params, tuple(offs), tuple(sizes), total)      # optim.FlatAdam lays its buffers out the same way
        self.grad_buckets = [(s, e, ev) for (s, e), ev in zip(bucket_ranges, events) if e > s]
        # gradients land in .grad here (autograd's accumulation semantics: a parameter that already holds a gradient adds to it)
        for prm, g in zip(params, grads):
            if prm.grad is None:
                prm.grad = g
            else:
                prm.grad = prm.grad + g
        if self.grad_ready_callback is not None:
            self.grad_ready_callback(self)       # e.g. dp.GradientAllReduce: bucket all-reduces on a side stream, overlapping
        return grads

    def bind_flat_grads(self):
        """Point every trainable parameter's ``.grad`` at its slice of ``flat_grad`` (the buffer the last backward wrote) —
        what the backward itself does when the parameter held no gradient; kept for callers that replaced ``.grad`` since."""
        params, grads = self._flat_views
        for prm, g in zip(params, grads):
            prm.grad = g

    # -- debugging / test helpers ----------------------------------------------------------------------
    def workspace_tensor(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """View of a named intermediate of the last forward (see enum tggcn_buf_id)."""
        dims, ws, _ = self._last
        off, nbytes = abi.workspace_view(dims, name)
        return ws[off:off + nbytes].view(dtype)[:math.prod(shape)].view(*shape)

    def check_persistent_kernels(self):
        """Wait for every call queued so far and raise if one of them reported a grid-barrier time-out or an fp16-split range
        violation.  forward() / backward() run the same test on their own for calls that have already finished; this is the
        synchronous form (end of an epoch, end of predict.py's loop, tests)."""
        self._poll_status(wait=True)
        dims, ws, _ = self._last
        stream = torch.cuda.current_stream(ws.device).cuda_stream
        abi.check(abi.lib().tggcn_sync_status(C.byref(dims), ws.data_ptr(), C.c_void_p(stream)), 'persistent kernels')


def select_model(model_name: str):
    """Same lookup as vhoi/models.py:1589-1595 for the model this package provides."""
    if model_name != '2G-GCN':
        raise KeyError(f'{model_name}: only the 2G-GCN model is provided by the B200 path')
    return TGGCN


def install_dropin():
    """Make the unchanged reference scripts (train.py:27, predict.py:36) pick up this class:
    ``vhoi.models.select_model('2G-GCN')`` builds its table from the module-global ``TGGCN``."""
    import vhoi.models as ref_models
    ref_models.TGGCN = TGGCN
    return ref_models
