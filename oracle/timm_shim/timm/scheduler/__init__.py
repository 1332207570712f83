def create_scheduler(args, optimizer):
    raise NotImplementedError("outer-loop LR scheduling is out of scope (SURVEY.md 2.1)")
