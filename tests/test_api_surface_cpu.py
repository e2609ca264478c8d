"""The public names a user of the reference expects resolve in this package."""
import importlib

import pytest

CHECK = {
 "triton_dist.utils": "initialize_distributed finalize_distributed nvshmem_create_tensor nvshmem_create_tensors nvshmem_free_tensor_sync nvshmem_barrier_all_on_stream NVSHMEM_SIGNAL_DTYPE dist_print rand_tensor get_bool_env get_int_env sleep_async has_tma is_nvshmem_multimem_supported supports_p2p_native_atomic launch_cooperative_grid_options get_triton_dist_world get_triton_dist_local_world_size LazyAllocator NVSHMEMLazyAllocator",
 "triton_dist.language": "wait consume_token notify symm_at rank num_ranks simt_exec_region",
 "triton_dist.kernels.nvidia": "ag_gemm create_ag_gemm_context gemm_persistent gemm_non_persistent gemm_rs create_gemm_rs_context create_gemm_ar_context create_ll_gemm_ar_context gemm_allreduce_op low_latency_gemm_allreduce_op ag_group_gemm create_ag_group_gemm_context create_moe_rs_context run_moe_reduce_rs fast_allgather create_fast_allgather_context fast_all_to_all create_all_to_all_context all_to_all_post_process create_ep_ll_a2a_ctx dispatch_kernel_v2 combine_kernel_v2 fused_sp_ag_attn_intra_node fused_sp_ag_attn_inter_node create_sp_ag_attention_context_intra_node gqa_fwd_batch_decode gqa_fwd_batch_decode_persistent gqa_fwd_batch_decode_intra_rank all_to_all_single_2d create_all_to_all_single_2d_context all_to_all_single_gemm create_all_to_all_single_gemm_context create_ulysses_sp_pre_attn_comm_context chunk_gated_delta_rule_fwd get_auto_all_gather_method AllGatherMethod cp_engine_producer_all_gather_intra_node cp_engine_producer_all_gather_inter_node moe_grouped_gemm moe_grouped_gemm_2weights transposed_moe_grouped_gemm run_moe_reduce_ar create_moe_ar_context ep_dispatch_token_inplace ep_combine_token_inplace bincount all_to_all_vdev_2d all_to_all_vdev_2d_offset pre_attn_qkv_pack_a2a_op qkv_bsnd_to_bnsd ulysses_sp_infer_gemm_a2a_op copy_tensor fill_tensor reduce_tensor swiglu_forward swiglu_backward matmul reduce_scatter_2d_op create_reduce_scater_2d_ctx ring_reduce calc_gather_scatter_index_triton histogram_by_expert_triton reduce_topk_tma get_tensorcore_tflops estimate_gemm_sol_time_ms estimate_reduce_scatter_time_ms estimate_all_gather_time_ms SpUlysessQKVGemmAll2AllKernel SpUlysessOAll2AllGemmKernel mega_kernel_dispatch_token_moe_grouped_gemm mega_kernel_moe_grouped_gemm_combine_token",
 "triton_dist.kernels.nvidia.allreduce": "create_allreduce_ctx all_reduce get_auto_allreduce_method",
 "triton_dist.kernels.allreduce": "AllReduceMethod OverlappingAllReduceMethod to_allreduce_method get_allreduce_methods",
 "triton_dist.function.nvidia": "TritonDistFusedEpMoeFunction MegaEpMoeFunction mega_ep_moe_autograd fused_ep_moe init_triton_dist_ep_op deinit_triton_dist_ep_op init_triton_dist_ep_ctx triton_dist_ep_op_initialized get_ep_capacity get_triton_dist_ep_stream get_triton_dist_ep_op TritonDistEpContext MoEOptimConfig get_moe_optim_config set_triton_dist_moe_profile_enabled get_triton_dist_moe_profile_enabled get_triton_dist_profile_output_dir custom_fwd custom_bwd",
 "triton_dist.layers.nvidia": "TP_MLP TP_Attn TP_MoE EP_MoE EPAll2AllLayer EPLowLatencyAllToAllLayer EpAll2AllFusedOp GemmARLayer AllGatherLayer SpGQAFlashDecodeAttention UlyssesSPAllToAllLayer CommOp PPCommLayer",
 "triton_dist.models": "ModelConfig AutoLLM AutoTokenizer DenseLLM Qwen3MoE KV_Cache Engine",
 "triton_dist.tune": "autotune",
 "triton_dist.autotuner": "contextual_autotune",
 "triton_dist.profiler_utils": "group_profile perf_func perf_func_with_l2_reset",
 "triton_dist.tools.profiler": "Profiler ProfilerBuffer alloc_profiler_buffer reset_profiler_buffer export_to_perfetto_trace",
 "triton_dist.test.utils": "assert_allclose assert_bitwise_equal LAYER_CONFIGS",
}


@pytest.mark.parametrize("mod", sorted(CHECK))
def test_api_surface(mod):
    m = importlib.import_module(mod)
    missing = [n for n in CHECK[mod].split() if not hasattr(m, n)]
    assert not missing, f"{mod} lacks {missing}"


LAYER_METHODS = {
    "TP_MLP": "_init_parameters _init_ctx torch_fwd dist_triton_fwd dist_triton_AR_fwd dist_triton_gemm_ar_fwd torch_ag_gemm dist_triton_ag_gemm torch_gemm_rs dist_triton_gemm_rs",
    "TP_Attn": "_init_parameters _init_ctx torch_fwd dist_triton_fwd dist_triton_AR_fwd dist_triton_gemm_ar_fwd",
    "TP_MoE": "torch_fwd dist_triton_fwd",
    "EP_MoE": "torch_fwd dist_triton_fwd",
    "EPAll2AllLayer": "preprocess dispatch combine dispatch_postprocess",
    "EPLowLatencyAllToAllLayer": "dispatch combine dump_dispatch_trace dump_combine_trace",
    "EpAll2AllFusedOp": "preprocess mega_preprocess_group_gemm mega_dispatch_group_gemm mega_group_gemm_combine get_nvshmem_size materialize",
    "GemmARLayer": "forward",
    "AllGatherLayer": "forward_pull forward_push_2d forward_push_3d forward_push_2d_ll forward_push_numa_2d forward_push_2d_ll_multimem",
    "SpGQAFlashDecodeAttention": "forward",
    "UlyssesSPAllToAllLayer": "pre_attn_qkv_pack_a2a",
    "CommOp": "read write set_signal wait_signal",
    "PPCommLayer": "send recv",
}


@pytest.mark.parametrize("cls", sorted(LAYER_METHODS))
def test_layer_methods(cls):
    """Method names of the reference's layer classes."""
    import triton_dist.layers.nvidia as L
    c = getattr(L, cls)
    missing = [n for n in LAYER_METHODS[cls].split() if not hasattr(c, n)]
    assert not missing, f"{cls} lacks {missing}"


def test_reference_module_paths_resolve():
    """Code written against the reference imports per-op modules (``triton_dist.kernels.nvidia.allgather_gemm`` ...): one meta-path
    finder (triton_dist/_module_map.py) maps every such path onto this package's modules; modules that exist on disk are untouched."""
    import importlib
    import triton_dist
    from triton_dist import _module_map
    from triton_dist.kernels.nvidia.allgather_gemm import ag_gemm, create_ag_gemm_context  # noqa: F401
    from triton_dist.kernels.nvidia.gemm_reduce_scatter import create_gemm_rs_context, gemm_rs  # noqa: F401
    from triton_dist.kernels.nvidia.low_latency_all_to_all_v2 import combine_kernel_v2, create_ep_ll_a2a_ctx, dispatch_kernel_v2  # noqa: F401
    from triton_dist.kernels.nvidia.moe_reduce_rs import create_moe_rs_context, run_moe_reduce_rs  # noqa: F401
    from triton_dist.language.extra import libshmem_device
    from triton_dist.layers.nvidia.ep_a2a_fused_layer import EpAll2AllFusedOp  # noqa: F401
    from triton_dist.layers.nvidia.tp_mlp import TP_MLP
    from triton_dist.mega_triton_kernel.models.dense import MegaDenseModel  # noqa: F401
    assert ag_gemm.__module__ == "triton_dist.ops.ag_gemm" and TP_MLP.__module__ == "triton_dist.parallel.tp_mlp"
    assert callable(libshmem_device.putmem_signal) and callable(libshmem_device.fcollect)
    import triton_dist.kernels.nvidia.allreduce as ar
    assert ar.__file__.endswith("allreduce.py")                       # a real module is not shadowed
    for name in _module_map.MODULE_MAP:                               # every mapped path imports and lists names
        m = importlib.import_module(name)
        assert dir(m), name
    with pytest.raises(ImportError):
        from triton_dist.kernels.nvidia.allgather_gemm import does_not_exist  # noqa: F401
    assert triton_dist.__version__


# SURVEY.md section 2.4 / 2.5 names every reference source file with its public entry points: each of them must resolve under the
# reference's own module path (triton_dist/_module_map.py maps those paths onto this package's subsystems)
K='triton_dist.kernels.nvidia.'
L='triton_dist.layers.nvidia.'
PER_FILE = {
 K+'allgather': 'AllGatherMethod get_auto_all_gather_method cp_engine_producer_all_gather_intra_node cp_engine_producer_all_gather_inter_node',
 K+'allgather_gemm': 'create_ag_gemm_context ag_gemm gemm_persistent gemm_non_persistent',
 K+'ag_gemm_threadblock_swizzle': 'threadblock_swizzle_allgather_gemm_kernel',
 K+'reduce_scatter': 'ReduceScatter2DContext create_reduce_scater_2d_ctx reduce_scatter_2d_op ring_reduce',
 K+'gemm_reduce_scatter': 'create_gemm_rs_context gemm_rs',
 K+'gemm_rs_threadblock_swizzle': 'threadblock_swizzle_gemm_reduce_scatter_kernel',
 'triton_dist.kernels.allreduce': 'AllReduceMethod OverlappingAllReduceMethod to_allreduce_method get_auto_all_reduce_method',
 K+'allreduce': 'create_allreduce_ctx all_reduce get_auto_allreduce_method',
 K+'gemm_allreduce': 'create_gemm_ar_context create_ll_gemm_ar_context gemm_allreduce_op low_latency_gemm_allreduce_op gemm_op allreduce_op',
 K+'group_gemm': 'moe_grouped_gemm moe_grouped_gemm_2weights transposed_moe_grouped_gemm',
 K+'moe_utils': 'calc_gather_scatter_index_triton calc_gather_scatter_index_v2_triton histogram_by_expert_triton reduce_topk_tma reduce_topk_non_tma',
 K+'allgather_group_gemm': 'create_ag_group_gemm_context ag_group_gemm',
 K+'moe_reduce_rs': 'create_moe_rs_context run_moe_reduce_rs run_moe_reduce_rs_triton_non_overlap',
 K+'moe_reduce_ar': 'create_moe_ar_context run_moe_reduce_ar',
 K+'low_latency_all_to_all': 'create_all_to_all_context fast_all_to_all all_to_all_post_process',
 K+'low_latency_all_to_all_v2': 'create_ep_ll_a2a_ctx LowlatencyDispatchContext LowlatencyCombineContext',
 K+'ep_a2a': 'ep_dispatch_token_inplace ep_combine_token_inplace get_ag_splits_and_recv_offset_for_dispatch bincount get_dispatch_send_reqs',
 K+'ep_a2a_intra_node': 'get_ag_splits_and_recv_offset_for_dispatch_intra_node',
 K+'ep_all2all_fused': 'mega_kernel_dispatch_token_moe_grouped_gemm mega_kernel_moe_grouped_gemm_combine_token get_ag_splits_and_recv_offset_for_dispatch',
 K+'all_to_all_vdev_2d_offset': 'create_context all_to_all_vdev_2d all_to_all_vdev_2d_offset all_to_all_v_offset_op all_to_all_v_offset_op_v2',
 K+'all_to_all_single_2d': 'create_all_to_all_single_2d_context all_to_all_single_2d',
 K+'all_to_all_single_gemm': 'create_all_to_all_single_gemm_context all_to_all_single_gemm gemm_only',
 K+'sp_ag_attention_intra_node': 'create_sp_ag_attention_context_intra_node fused_sp_ag_attn_intra_node',
 K+'sp_ag_attention_inter_node': 'create_sp_ag_attention_context_inter_node fused_sp_ag_attn_inter_node',
 K+'flash_decode': 'gqa_fwd_batch_decode gqa_fwd_batch_decode_persistent gqa_fwd_batch_decode_intra_rank gqa_fwd_batch_decode_aot gqa_fwd_batch_decode_persistent_aot gqa_fwd_batch_decode_intra_rank_aot',
 K+'low_latency_allgather': 'create_fast_allgather_context fast_allgather',
 K+'ulysses_sp_dispatch': 'create_ulysses_sp_pre_attn_comm_context pre_attn_qkv_pack_a2a_op qkv_bsnd_to_bnsd',
 K+'sp_ulysess_qkv_gemm_all2all': 'SpUlysessQKVGemmAll2AllKernel',
 K+'sp_ulysess_o_all2all_gemm': 'SpUlysessOAll2AllGemmKernel',
 K+'ulysses_sp_infer_gemm_a2a': 'UlyssesSpInferPreAttnContext ulysses_sp_infer_gemm_a2a_op pre_attn_a2a_comm_only',
 K+'p2p': 'p2p_set_signal p2p_wait_signal p2p_copy_kernel p2p_put_kernel p2p_copy_remote_to_local_kernel',
 K+'memory_ops': 'copy_tensor fill_tensor reduce_tensor',
 K+'swiglu': 'swiglu_forward swiglu_backward',
 K+'gdn': 'chunk_gated_delta_rule_fwd',
 K+'gemm': 'get_config_space matmul matmul_tma matmul_persistent matmul_tma_persistent matmul_descriptor_persistent',
 K+'gemm_perf_model': 'get_tensorcore_tflops get_dram_gbps estimate_gemm_sol_time_ms',
 K+'comm_perf_model': 'estimate_reduce_scatter_time_ms estimate_all_gather_time_ms get_nic_gbps_per_gpu',
 L+'tp_mlp': 'TP_MLP', L+'tp_attn': 'TP_Attn', L+'tp_moe': 'TP_MoE', L+'ep_moe': 'EP_MoE',
 L+'ep_a2a_layer': 'EPConfig DispatchCombineContext EPAll2AllLayer',
 L+'ep_ll_a2a_layer': 'EPLowLatencyAllToAllLayer',
 L+'ep_a2a_fused_layer': 'EpAll2AllFusedOp',
 L+'gemm_allreduce_layer': 'GemmARLayer',
 L+'low_latency_allgather_layer': 'AllGatherLayer',
 L+'sp_flash_decode_layer': 'SpGQAFlashDecodeAttention',
 L+'ulysses_sp_a2a_layer': 'UlyssesSPAllToAllLayer',
 L+'p2p': 'CommOp', L+'pp_block': 'PPCommLayer PyTorchP2P',
 'triton_dist.function.nvidia.ep_moe_fused': 'TritonDistFusedEpMoeFunction',
 'triton_dist.function.nvidia.common': 'init_triton_dist_ep_op MoEOptimConfig',
 'triton_dist.mega_triton_kernel': 'ModelBuilder',
 'triton_dist.mega_triton_kernel.models': 'DenseModel',
}


@pytest.mark.parametrize("mod", sorted(PER_FILE))
def test_reference_file_entry_points(mod):
    m = importlib.import_module(mod)
    missing = [n for n in PER_FILE[mod].split() if not hasattr(m, n)]
    assert not missing, f"{mod} lacks {missing}"


# SURVEY.md sections 2.2 / 2.3 / 2.7 / 2.9: device language, NVSHMEM-style device API, runtime helpers, megakernel sub-modules, tools
LANGUAGE_AND_RUNTIME = {
 'triton_dist.language': 'wait consume_token notify symm_at rank num_ranks simt_exec_region vector make_vector zeros_vector extern_call',
 'triton_dist.language.extra.cuda.language_extra': '__syncthreads __fence tma_sync multimem_st_b32 multimem_st_b64 multimem_st_v2 multimem_st_v4 multimem_ld_reduce_v4 tid ntid laneid red_release arrive_inc ld ld_vector st_vector ld_acquire st atomic_add atomic_add_per_warp wait_eq __shfl_sync_i32 __shfl_up_sync_i32 __shfl_down_sync_i32 __ballot_sync atomic_cas globaltimer smid membar fence pack_b32_v2 pack unpack',
 'triton_dist.language.extra.cuda.libnvshmem_device': 'NVSHMEM_CMP_EQ NVSHMEM_CMP_GE NVSHMEM_SIGNAL_SET NVSHMEM_SIGNAL_ADD my_pe n_pes team_my_pe team_n_pes int_p remote_ptr remote_mc_ptr barrier barrier_block barrier_warp barrier_all barrier_all_block barrier_all_warp sync_all sync_all_block sync_all_warp team_sync_block team_sync_warp quiet fence getmem getmem_nbi getmem_warp getmem_block getmem_nbi_block putmem putmem_nbi putmem_warp putmem_block putmem_nbi_block putmem_signal putmem_signal_nbi putmem_signal_block putmem_signal_nbi_block putmem_signal_warp signal_op signal_wait_until broadcast broadcast_warp broadcast_block broadcastmem_block fcollect fcollect_block team_translate_pe',
 'triton_dist.language.extra.libshmem_device': 'my_pe n_pes putmem_block getmem_block signal_op signal_wait_until barrier_all',
 'triton_dist.kernels.nvidia.common_ops': 'unsafe_barrier_on_this_grid cooperative_barrier_on_this_grid barrier_all_intra_node_atomic_cas_block barrier_all_intra_node_non_atomic barrier_all_intra_node_non_atomic_block _wait_eq_cuda _set_signal_cuda _memcpy_async_cuda bisect_left bisect_right bisect_left_aligned bisect_right_aligned BarrierAllContext barrier_all_on_stream',
 'triton_dist.utils': 'wait_until_max_gpu_clock_or_warning triton_dist_key barrier_async generate_data _make_tensor init_seed has_fullmesh_nvlink get_nvlink_max_speed_gbps get_intranode_max_speed_gbps LazyTensor LazyAllocator NVSHMEMLazyAllocator nvshmem_create_tensors',
 'triton_dist.nv_utils': 'has_fullmesh_nvlink get_nvlink_max_speed_gbps get_intranode_max_speed_gbps ensure_nvml_initialized with_pynvml nvsmi get_max_gpu_clock_rate_in_khz get_current_gpu_clock_rate_in_khz get_nvlink_adjacency_matrix has_fullmesh_nvlink_pynvml calculate_pcie_bandwidth_gbps get_pcie_link_max_speed_gbps get_numa_node get_device_name gpu_uuid_string get_physical_gpu_uuid get_physical_device_count is_gpu_max_performance_mode get_nvlink get_nvcc',
 'triton_dist.jit': 'jit',
 'triton_dist.tools.compile_aot': 'aot_compile_spaces',
 'triton_dist.tune': 'autotune',
 'triton_dist.autotuner': 'contextual_autotune ContextualAutoTuner',
 'triton_dist.tools.profiler': 'Profiler ProfilerBuffer alloc_profiler_buffer reset_profiler_buffer export_to_perfetto_trace',
 'triton_dist.tools.profiler.language': 'Profiler',
 'triton_dist.tools.profiler.viewer': 'export_to_perfetto_trace',
 'triton_dist.tools.tune.tune_gemm': 'main',
 'triton_dist.tools.tune.find_topk': 'main',
 'triton_dist.benchmark.bench_allgather_gemm': '', 'triton_dist.benchmark.bench_tp_mlp': '', 'triton_dist.benchmark.bench_tp_attn': '', 'triton_dist.benchmark.bench_pp': '',
 'triton_dist.models.engine': 'Engine', 'triton_dist.models.utils': 'logger seed_everything sample_token init_model_cpu',
 'triton_dist.models.kv_cache': 'KV_Cache', 'triton_dist.models.config': 'ModelConfig', 'triton_dist.models.dense': 'DenseLLM DenseLLMLayer', 'triton_dist.models.qwen_moe': 'Qwen3MoE',
 'triton_dist.mega_triton_kernel.core.scheduler': '', 'triton_dist.mega_triton_kernel.core.code_generator': '', 'triton_dist.mega_triton_kernel.core.registry': '', 'triton_dist.mega_triton_kernel.core.config': '',
 'triton_dist.mega_triton_kernel.tasks.flash_attn': '', 'triton_dist.mega_triton_kernel.tasks.prefetch': '', 'triton_dist.mega_triton_kernel.tasks.barrier':'', 'triton_dist.mega_triton_kernel.tasks.elementwise':'', 'triton_dist.mega_triton_kernel.tasks.activation':'',
 'triton_dist.mega_triton_kernel.kernels.flash_attn': '', 'triton_dist.mega_triton_kernel.kernels.linear': '',
 'triton_dist.mega_triton_kernel.test.models.model_server': '', 'triton_dist.mega_triton_kernel.test.models.chat': '',
}


@pytest.mark.parametrize("mod", sorted(LANGUAGE_AND_RUNTIME))
def test_language_runtime_and_tool_names(mod):
    m = importlib.import_module(mod)
    missing = [n for n in LANGUAGE_AND_RUNTIME[mod].split() if not hasattr(m, n)]
    assert not missing, f"{mod} lacks {missing}"

