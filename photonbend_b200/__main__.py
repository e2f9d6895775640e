from photonbend_b200.scripts.main import main

if __name__ == "__main__":
    main()
