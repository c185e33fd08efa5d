from os.path import dirname

print()
print("fitnoise_b200 is intended to be used as a library, from Python (import fitnoise_b200 as fitnoise)")
print("or from R through the fitnoise R package files shipped in", dirname(__file__) + "/R")
