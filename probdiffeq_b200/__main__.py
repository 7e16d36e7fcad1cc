from probdiffeq_b200.build import build

if __name__ == "__main__":
    print(build())
