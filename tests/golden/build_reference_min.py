# minimal in-place build of the three hot-path Cython extensions of the reference copy
from setuptools import setup, Extension
from Cython.Build import cythonize
from numpy import get_include
inc = get_include()
exts = [
 Extension('bhmm.hidden.impl_c.hidden', ['bhmm/hidden/impl_c/hidden.pyx', 'bhmm/hidden/impl_c/_hidden.c'], include_dirs=['bhmm/hidden/impl_c', inc]),
 Extension('bhmm.output_models.impl_c.discrete', ['bhmm/output_models/impl_c/discrete.pyx', 'bhmm/output_models/impl_c/_discrete.c'], include_dirs=['bhmm/output_models/impl_c', inc]),
 Extension('bhmm.output_models.impl_c.gaussian', ['bhmm/output_models/impl_c/gaussian.pyx', 'bhmm/output_models/impl_c/_gaussian.c'], include_dirs=['bhmm/output_models/impl_c', inc]),
]
setup(name='bhmm_min', ext_modules=cythonize(exts, language_level=2), script_args=['build_ext', '--inplace'])
