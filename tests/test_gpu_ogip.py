"""Files -> FixedData -> likelihood on the GPU: the ingested OGIP data drive the kernels and
agree with the oracle evaluated on the same arrays (dense and CSR-by-channel responses)."""

import numpy as np
import pytest

import ogip_cases
from elisa_b200 import models as M
from elisa_b200.ogip import load_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp('ogip')
    return str(d), ogip_cases.build(str(d))


@pytest.mark.parametrize('case,stat', [('a_varlen_arf_back_grouped', 'wstat'), ('a_varlen_arf_back_grouped', 'pstat'),
                                       ('a_no_background_sparse', 'cstat'), ('b_fixed_rsp_rate_net', 'chi2'),
                                       ('c_type2_row2_gaussian_back', 'pgstat')])
def test_ingested_data_through_the_likelihood(lib, files, case, stat, monkeypatch):
    from parity import RTOL, compare
    from test_gpu_parity import _theta_for

    tmp, cases = files
    monkeypatch.chdir(tmp)
    d = load_data(**{k: v for k, v in cases[case].items() if not k.startswith('_')})
    model = M.CutoffPL(K=M.UniformParameter('K', 0.05, 1e-6, 1e3)) + M.Gauss(El=6.4, sigma=0.3)
    cfg = dict(data=[d], model=[model], stat=[stat], theta=_theta_for(model, 24, 9))
    err = compare(cfg)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, (case, stat, err)


def test_joint_fit_of_three_ingested_spectra(lib, files, monkeypatch):
    from parity import RTOL, compare
    from test_gpu_parity import _theta_for

    tmp, cases = files
    monkeypatch.chdir(tmp)
    names = ['a_varlen_arf_back_grouped', 'b_fixed_rsp_rate_net', 'c_type2_row2_gaussian_back']
    data = [load_data(**{k: v for k, v in cases[n].items() if not k.startswith('_')}) for n in names]
    model = M.PowerLaw(K=M.UniformParameter('K', 0.1, 1e-6, 1e3))
    f = M.Constant()
    cfg = dict(data=data, model=[model, f * model, model], stat=['wstat', 'chi2', 'pgstat'],
               theta=_theta_for((f * model), 16, 4)[:, [1, 2, 0]])
    err = compare(cfg)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err
