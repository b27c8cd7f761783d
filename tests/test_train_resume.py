"""CPU: the resume logic of tools/train_net.py (reference tools/train_net.py:79-105): model_final.pkl ends training, else the
newest model_iter<N>.pkl restarts at N + 1."""
import importlib.util
import os


def _mod():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('train_net', os.path.join(root, 'tools', 'train_net.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_find_resume_point(tmp_path):
    tn = _mod()
    d = str(tmp_path)
    assert tn.find_resume_point(os.path.join(d, 'missing')) == (0, None)
    assert tn.find_resume_point(d) == (0, None)
    for n in ('model_iter19999.pkl', 'model_iter39999.pkl', 'model_iter7.pkl', 'net.pbtxt'):
        open(os.path.join(d, n), 'wb').close()
    assert tn.find_resume_point(d) == (40000, os.path.join(d, 'model_iter39999.pkl'))
    open(os.path.join(d, 'model_final.pkl'), 'wb').close()
    assert tn.find_resume_point(d) == ('final', os.path.join(d, 'model_final.pkl'))
