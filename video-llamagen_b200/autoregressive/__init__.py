"""Import-path shim: ``autoregressive.models.{gpt,generate}`` resolve to the vlg drop-ins when
``video-llamagen_b200/`` precedes the reference checkout on sys.path.  Other sub-packages of the
reference (``autoregressive.sample`` ...) stay reachable: the reference tree, when given through
``$VLG_REFERENCE``, is appended to this package's search path."""
import os

_ref = os.environ.get("VLG_REFERENCE")
if _ref and os.path.isdir(os.path.join(_ref, "autoregressive")):
    __path__.append(os.path.join(_ref, "autoregressive"))
