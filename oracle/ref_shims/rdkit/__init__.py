"""TEST INFRASTRUCTURE ONLY — stub so ``notorch/types.py:5`` (``from rdkit.Chem import Mol``) imports."""
