"""CPU checks of the measurement contract: the bench line committed from the last GPU run carries every key the
driver reads, the roofline arithmetic is self-consistent, and bench.py's host-side helpers behave (no GPU here)."""
import json
import os

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = ['r1_j_bench_g3.json', 'r1_k_bench_short_with_torch_gpu_baseline.json']


@pytest.mark.parametrize('name', LINES)
def test_committed_bench_line_follows_the_contract(name):
    line = json.loads(open(os.path.join(REPO, 'profiles', name)).read().strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert key in line, key
    assert line['higher_is_better'] is True and line['scaling'] == 'weak' and line['vs_baseline'] is None
    assert 'workload' in line['config'] and 'model' not in line['config'] and 'l2' in line['config']
    assert line['warmup'] >= 3 and line['gpu_launches'] > 0 and line['data'] == 'synthetic'
    e2e = line['e2e']
    assert e2e['h2d_bytes_per_step'] > 0 and e2e['d2h_bytes_per_step'] > 0 and e2e['unit'] == line['unit']
    assert e2e['value'] != line['value']                       # measured separately, not a copy
    roof = line['roofline']
    assert roof['bound'] in ('hbm', 'tensor') and roof['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-9 and 0 < roof['frac'] < 1
    assert roof['traffic'] is None or roof['traffic'] > 0
    assert set(line['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    assert not set(line['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    # whole-job throughput is what the step time says it is
    q = line['queries_per_step_per_gpu'] * line['n_gpus']
    assert abs(line['value'] - q / (line['ms_per_step'] / 1e3)) / line['value'] < 1e-6
    if line['cpu_baseline'] is not None:
        assert set(line['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'}
        assert line['cpu_baseline']['kind'] in ('reference', 'port')
    if line.get('torch_gpu_baseline'):
        t = line['torch_gpu_baseline']
        assert t['tf32'] is False and t['o4d_over_torch_eager'] > 10.0 and t['max_rel_err_vs_torch_eager'] < 1e-3


@pytest.mark.parametrize('name', ['r2_j_bench.json', 'r2_k_bench_final.json'])
def test_round2_bench_lines_carry_the_round2_keys(name):
    """The full-round lines of round 2: the contract keys plus the real-reference baselines, config-3 / config-4 /
    config-5 keys and their internal consistency."""
    line = json.loads(open(os.path.join(REPO, 'profiles', name)).read().strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline',
                'torch_gpu_baseline', 'encoder', 'carla_config3', 'strong_scaling', 'train_step', 'kernel_families'):
        assert key in line, key
    assert line['dtype'] == 'bf16x3' and line['vs_baseline'] is None and line['warmup'] >= 3
    q = line['queries_per_step_per_gpu'] * line['n_gpus']
    assert abs(line['value'] - q / (line['ms_per_step'] / 1e3)) / line['value'] < 1e-6
    roof = line['roofline']
    assert roof['bound'] == 'tensor' and abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-9
    fam = line['kernel_families']
    assert roof['kernel'] in fam and sum(f['ms'] for f in fam.values()) < line['ms_per_step']
    assert abs(roof['share_of_step'] - fam[roof['kernel']]['ms'] / sum(f['ms'] for f in fam.values())) < 1e-6
    assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] >= 1
    t = line['torch_gpu_baseline']
    assert t['kind'] == 'reference' and t['tf32'] is False and t['warmup_passes'] == 10 and t['timed_passes'] == 20
    assert t['o4d_over_torch_eager'] > 10.0                                   # the north star's own target
    err = t['rel_err_o4d_vs_reference_all_queries']
    assert err['rows'] == 534528 and err['p9999'] < 1e-4 and err['rows_above_1e-3'] <= 4   # near-tie neighbour flips only
    assert line['e2e']['bit_identical_to_device_path'] is True
    enc = line['encoder']
    assert enc['ms'] < 8.0 and enc['critical_path']['ms'] < enc['ms'] and enc['reference']['gpu_eager']['kind'] == 'reference'
    c3 = line['carla_config3']
    assert c3['m_abstract'] == 2124 and c3['d_out'] == 18 and c3['max_rel_err_vs_reference_golden'] < 1e-3
    ss = line['strong_scaling']
    assert ss['queries'] == 2097152 and set(ss['by_batch']) == {'8192', '32768', '131072'}
    assert ss['sharded_equals_single_rank_bitwise'] is True and ss['max_abs_diff_sharded_vs_single'] == 0.0
    tr = line['train_step']
    assert tr['queries_per_sample'] == 4 * 17203 and tr['ms_per_step'] < 166.0 and 'bf16' in tr and 'roofline' in tr
    ref = tr['reference_autograd_decoder_frame']
    assert ref['kind'].startswith('reference') and ref['speedup_vs_reference_fp32'] > 3.0
    assert ref['abstract_feature_and_global_gradient_rel_l2_vs_reference_autograd'] < 1e-3


def test_reference_arm_line_follows_the_contract():
    line = json.loads(open(os.path.join(REPO, 'profiles', 'r1_j_bench_reference_cpu.json')).read().strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['gpu_launches'] == 0
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['cpu_baseline']['value'] == line['value'] and line['cpu_baseline']['cores'] >= 1


def test_bench_host_helpers():
    import bench
    cfg = bench.workload_config(32768)
    assert cfg['implicit_batch_size'] == 32768 and '524288' in cfg['workload']
    abstract, glob = bench.golden_scene()
    assert abstract.shape == (531, 291) and glob.shape == (128,) and abstract.dtype == torch.float32
    # 47.89 MFLOP per query is the figure roofline.achieved is built on
    assert abs(bench.FLOP_PER_QUERY - 47.89e6) / 47.89e6 < 2e-3
