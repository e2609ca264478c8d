"""Reference module paths -> the modules of this package.

The reference keeps one Python module per op / layer (``triton_dist.kernels.nvidia.allgather_gemm``, ``triton_dist.layers.nvidia.tp_mlp``,
``triton_dist.mega_triton_kernel.models.dense`` ...); this package groups the same functionality by subsystem (``ops``, ``parallel``,
``mega_kernel``, ``language``).  Code written against the reference imports by the old paths, so instead of ~70 re-export stubs one
meta-path finder serves them: ``import triton_dist.kernels.nvidia.allgather_gemm`` yields a facade module whose attributes resolve, in
order, in the implementing module(s) listed here and then in the package-level namespace of that family (``triton_dist.kernels.nvidia``
/ ``triton_dist.layers.nvidia``, which hold every public name of the reference's ``__init__``).  Real submodules always win: the finder
only answers for names that do not exist on disk.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys
import types

_K = "triton_dist.kernels.nvidia."
_L = "triton_dist.layers.nvidia."
_O = "triton_dist.ops."
_P = "triton_dist.parallel."

MODULE_MAP = {
    # kernels
    "triton_dist.kernels.common_ops": ["triton_dist.utils", _O + "comm", "triton_dist.lk.stdlib", _O + "compat"],
    _K + "common_ops": ["triton_dist.utils", _O + "comm", "triton_dist.lk.stdlib", _O + "compat"],
    _K + "allgather_gemm": [_O + "ag_gemm"],
    _K + "ag_gemm_threadblock_swizzle": [_O + "tile_swizzle", _O + "ag_gemm"],
    _K + "allgather": [_O + "allgather"],
    _K + "gemm_reduce_scatter": [_O + "gemm_rs"],
    _K + "gemm_rs_threadblock_swizzle": [_O + "tile_swizzle", _O + "gemm_rs"],
    _K + "reduce_scatter": [_O + "comm", _O + "gemm_rs"],
    _K + "gemm_allreduce": [_O + "gemm_ar"],
    _K + "gemm": [_O + "gemm"],
    _K + "group_gemm": [_O + "moe"],
    _K + "moe_utils": [_O + "moe"],
    _K + "allgather_group_gemm": [_O + "moe"],
    _K + "threadblock_swizzle_ag_moe": [_O + "tile_swizzle", _O + "moe"],
    _K + "threadblock_swizzle_ag_moe_triton": [_O + "tile_swizzle", _O + "moe"],
    _K + "moe_reduce_rs": [_O + "moe"],
    _K + "moe_reduce_ar": [_O + "moe"],
    _K + "low_latency_allgather": [_O + "comm"],
    _K + "low_latency_all_to_all": [_O + "all_to_all"],
    _K + "low_latency_all_to_all_v2": [_O + "ep_a2a"],
    _K + "ep_a2a": [_O + "ep_normal", _O + "compat"],
    _K + "ep_a2a_intra_node": [_O + "ep_normal", _O + "compat"],
    _K + "ep_all2all_fused": [_O + "ep_mega", _O + "compat"],
    _K + "all_to_all_single_2d": [_O + "all_to_all"],
    _K + "all_to_all_single_gemm": [_O + "compat", _O + "ag_gemm"],
    _K + "all_to_all_vdev_2d_offset": [_O + "compat", _O + "all_to_all"],
    _K + "all_to_all_vdev_2d_offset_inter_node": [_O + "compat", _O + "all_to_all"],
    _K + "sp_ag_attention_intra_node": [_P + "sp", _O + "flash_attn"],
    _K + "sp_ag_attention_inter_node": [_P + "sp", _O + "flash_attn"],
    _K + "sp_ulysess_qkv_gemm_all2all": [_O + "compat", _O + "gemm_a2a"],
    _K + "sp_ulysess_o_all2all_gemm": [_O + "compat"],
    _K + "ulysses_sp_dispatch": [_O + "compat", _P + "sp"],
    _K + "ulysses_sp_infer_gemm_a2a": [_O + "compat", _O + "gemm_a2a"],
    _K + "flash_decode": [_O + "flash_decode"],
    _K + "gdn": [_O + "gdn"],
    _K + "memory_ops": [_O + "comm"],
    _K + "p2p": [_O + "p2p"],
    _K + "swiglu": [_O + "compat", _O + "elementwise"],
    _K + "gemm_perf_model": [_O + "perf_model"],
    _K + "comm_perf_model": [_O + "perf_model"],
    # layers
    _L + "tp_mlp": [_P + "tp_mlp"],
    _L + "tp_attn": [_P + "tp_attn"],
    _L + "tp_moe": [_P + "tp_moe"],
    _L + "ep_moe": [_P + "ep"],
    _L + "ep_a2a_layer": [_P + "ep"],
    _L + "ep_ll_a2a_layer": [_P + "ep"],
    _L + "ep_a2a_fused_layer": [_P + "ep", _O + "ep_mega"],
    _L + "gemm_allreduce_layer": [_P + "misc"],
    _L + "low_latency_allgather_layer": [_P + "misc"],
    _L + "p2p": [_P + "pp"],
    _L + "pp_block": [_P + "pp"],
    _L + "sp_flash_decode_layer": [_P + "sp"],
    _L + "ulysses_sp_a2a_layer": [_P + "sp"],
    # device language
    "triton_dist.language.core": ["triton_dist.language"],
    "triton_dist.language.distributed_ops": ["triton_dist.language"],
    "triton_dist.language.simt_ops": ["triton_dist.language"],
    "triton_dist.language.extra": ["triton_dist.language.shmem", "triton_dist.language"],
    "triton_dist.language.extra.libshmem_device": ["triton_dist.lk.shmem", "triton_dist.language.shmem"],
    "triton_dist.language.extra.language_extra": ["triton_dist.lk.language_extra", "triton_dist.language", "triton_dist.language.shmem"],
    "triton_dist.language.extra.utils": ["triton_dist.language"],
    "triton_dist.language.extra.cuda": ["triton_dist.language.shmem", "triton_dist.language"],
    "triton_dist.language.extra.cuda.language_extra": ["triton_dist.lk.language_extra", "triton_dist.language", "triton_dist.language.shmem"],
    "triton_dist.language.extra.cuda.libnvshmem_device": ["triton_dist.lk.shmem", "triton_dist.language.shmem"],
    # tools / misc
    "triton_dist.nv_utils": ["triton_dist.utils.topology", "triton_dist.utils", "triton_dist._build"],
    "triton_dist.tools.compile": ["triton_dist.tools.compile_aot"],
    "triton_dist.tools.compile.compile": ["triton_dist.tools.compile_aot"],
    "triton_dist.tools.profiler.context": ["triton_dist.tools.profiler"],
    "triton_dist.tools.profiler.language": ["triton_dist.tools.profiler"],
    "triton_dist.tools.profiler.viewer": ["triton_dist.tools.profiler"],
    # megakernel
    "triton_dist.mega_triton_kernel": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.builder": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.scheduler": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.task_base": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.graph": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.models": ["triton_dist.mega_kernel.dense", "triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.models.dense": ["triton_dist.mega_kernel.dense"],
    "triton_dist.mega_triton_kernel.models.paged_kv_cache": ["triton_dist.models.paged_kv_cache"],
    "triton_dist.mega_triton_kernel.models.model_builder": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.models.utils": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.task_context": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.allreduce": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.flash_decode": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.norm": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.linear": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.code_generator": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.registry": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.config": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.core.utils": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.flash_attn": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.prefetch": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.barrier": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.elementwise": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.activation": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.tasks.utils": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.flash_attn": ["triton_dist.lk.kernels.flash_mma", "triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.flash_decode": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.linear": ["triton_dist.lk.kernels.linear_mma", "triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.norm": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.allreduce": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.activation": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.elementwise": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.prefetch": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.barrier": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.mlp_fc1": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.kernels.utils": ["triton_dist.mega_kernel"],
    "triton_dist.mega_triton_kernel.test": ["triton_dist.mega_kernel.server"],
    "triton_dist.mega_triton_kernel.test.models": ["triton_dist.mega_kernel.server"],
    "triton_dist.mega_triton_kernel.test.models.model_server": ["triton_dist.mega_kernel.server"],
    "triton_dist.mega_triton_kernel.test.models.chat": ["triton_dist.mega_kernel.server"],
}
_FAMILY = {_K: "triton_dist.kernels.nvidia", _L: "triton_dist.layers.nvidia"}


class _Facade(types.ModuleType):
    """A module whose attributes are looked up in the implementing modules (imported on first use)."""

    def __init__(self, name, targets):
        super().__init__(name, f"Facade for the reference module path {name!r}: resolves names in {', '.join(targets)}.")
        self.__dict__["_td_targets"] = list(targets)
        self.__path__ = []          # importable as a package so that mapped submodules resolve

    def __getattr__(self, attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        for t in self.__dict__["_td_targets"]:
            m = importlib.import_module(t)
            if hasattr(m, attr):
                v = getattr(m, attr)
                self.__dict__[attr] = v
                return v
        raise AttributeError(f"module {self.__name__!r} (reference path) has no attribute {attr!r}; looked in {self.__dict__['_td_targets']}")

    def __dir__(self):
        names = set()
        for t in self.__dict__["_td_targets"]:
            names.update(n for n in dir(importlib.import_module(t)) if not n.startswith("_"))
        return sorted(names)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in MODULE_MAP:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        targets = list(MODULE_MAP[spec.name])
        for prefix, family in _FAMILY.items():
            if spec.name.startswith(prefix):
                targets.append(family)
                if prefix == _K and _O + "compat" not in targets:
                    targets.append(_O + "compat")          # reference spellings that are thin wrappers / aliases (incl. lazily resolved ones)
        return _Facade(spec.name, targets)

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())          # appended: modules that exist on disk are found first
