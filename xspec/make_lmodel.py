import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lineprofiles_b200.xspec_models import lmodel_text
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lmodel.dat"), "w").write(lmodel_text())
