"""Result files in the reference's wire format, so that its R post-processing
(`r/paper/*.R`) reads the output of chains that ran on the B200 estimator.

The reference stores four gzip-compressed JSON documents per simulation under
`<output_path>/<sim_name>/` (helpers/file_system.py:40-75, parameter/mcmc/output.py:147-222):
`mcmc_output.json.gz`, `data.json.gz`, `settings.json.gz`, `description.txt.gz`; NumPy arrays are
written as nested lists, burn-in iterations are dropped, and every document repeats the
simulation name / description / time."""
import copy
import gzip
import json
import os
import time

import numpy as np


def _jsonable(value):
    if isinstance(value, np.ndarray):
        return value.tolist()
    if isinstance(value, (np.floating, np.integer, np.bool_)):
        return value.item()
    if isinstance(value, dict):
        return {str(k): _jsonable(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_jsonable(v) for v in value]
    return value


def write_to_json(data, output_path, sim_name, output_type, as_gzip=True):
    """file_system.py:40-75: `<output_path>/<sim_name>/<output_type>[.gz]`, UTF-8 JSON."""
    file_name = os.path.join(output_path, sim_name, output_type)
    os.makedirs(os.path.dirname(file_name), exist_ok=True)
    payload = json.dumps({str(k): _jsonable(v) for k, v in data.items()})
    if as_gzip:
        with gzip.GzipFile(file_name + '.gz', 'w') as fout:
            fout.write(payload.encode('utf-8'))
    else:
        with open(file_name, 'w') as fout:
            fout.write(payload)
    return file_name + ('.gz' if as_gzip else '')


def read_json(file_name):
    opener = gzip.open if file_name.endswith('.gz') else open
    with opener(file_name, 'rt', encoding='utf-8') as fin:
        return json.load(fin)


def compile_results(chain, model, settings, sampler_name, sim_name=None, sim_desc=None):
    """output.py:147-190 for a chain result dict (keys of `ParticleMH.run` / `LockstepChains.run`:
    params, prop_params, accepted, accept_prob, states, nat_gradient, hess -- whatever is present is
    cut to the post-burn-in iterations, the rest is stored as null)."""
    no_iters = int(settings['no_iters'])
    burn = int(settings.get('no_burnin_iters', 0))
    idx = slice(burn, no_iters)
    now = time.strftime("%c")

    def cut(key):
        value = chain.get(key)
        return None if value is None else np.asarray(value)[idx]

    mcmcout = {k: cut(k) for k in ('params', 'prop_params', 'states', 'accept_prob', 'accepted',
                                   'no_samples_hess_est', 'nat_gradient', 'hess')}
    mcmcout.update({'no_hessians_corrected': chain.get('no_hessians_corrected', 0),
                    'iter_hessians_corrected': chain.get('iter_hessians_corrected', []),
                    'simulation_description': sim_desc, 'simulation_name': sim_name,
                    'simulation_time': now, 'time_per_iteration': chain.get('time_per_iteration')})
    data = {'observations': np.asarray(model.obs), 'simulation_description': sim_desc,
            'simulation_name': sim_name, 'simulation_time': now}
    if getattr(model, 'states', None) is not None:
        data['states'] = np.asarray(model.states)
    out_settings = copy.deepcopy(dict(settings))
    out_settings.update({'sampler_name': sampler_name, 'simulation_description': sim_desc,
                         'simulation_name': sim_name, 'simulation_time': now})
    return mcmcout, data, out_settings


def store_results_to_file(chain, model, settings, sampler_name, output_path, sim_name, sim_desc=None):
    """output.py:192-222: writes the four documents; returns their paths."""
    if output_path is None:
        raise ValueError("No output path given...")
    mcout, data, out_settings = compile_results(chain, model, settings, sampler_name, sim_name, sim_desc)
    desc = {'description': out_settings['simulation_description'], 'time': out_settings['simulation_time']}
    return [write_to_json(mcout, output_path, sim_name, 'mcmc_output.json'),
            write_to_json(data, output_path, sim_name, 'data.json'),
            write_to_json(out_settings, output_path, sim_name, 'settings.json'),
            write_to_json(desc, output_path, sim_name, 'description.txt')]
