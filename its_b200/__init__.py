"""Importable alias of the package directory
`inference-time-scaling-for-diffusion-models-beyond-scaling-denoising-steps_b200/`
(hyphens are not legal in a Python module name).  `import its_b200` executes that
package's __init__ with this module's __path__ pointing at it, so
`its_b200.Diffusion.Model` etc. resolve to the files in the hyphenated directory."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "inference-time-scaling-for-diffusion-models-beyond-scaling-denoising-steps_b200")
__path__ = [_PKG]
with open(_os.path.join(_PKG, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG, "__init__.py"), "exec"))
del _f
